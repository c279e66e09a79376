// Test infrastructure only (oracle/): stand-in for Catch2 3.8.0 (reference CMakeLists.txt:14-19)
// so that /root/reference/tests/unit_tests.cpp compiles UNMODIFIED.  Self-registering test cases,
// REQUIRE that records a failure and leaves the case, and a main() that runs everything.
#pragma once
#include <cstdio>
#include <exception>
#include <vector>

namespace catch_shim {
struct Case {
    const char* name;
    void (*fn)();
};
inline std::vector<Case>& registry() {
    static std::vector<Case> r;
    return r;
}
inline int& failures() {
    static int f = 0;
    return f;
}
struct Registrar {
    Registrar(const char* name, void (*fn)()) { registry().push_back({name, fn}); }
};
struct RequireFailed {};
} // namespace catch_shim

#define CATCH_SHIM_CAT2(a, b) a##b
#define CATCH_SHIM_CAT(a, b)  CATCH_SHIM_CAT2(a, b)

#define TEST_CASE(name, tags)                                                                    \
    static void                   CATCH_SHIM_CAT(catch_shim_case_, __LINE__)();                  \
    static catch_shim::Registrar  CATCH_SHIM_CAT(catch_shim_reg_, __LINE__)(name,                \
        &CATCH_SHIM_CAT(catch_shim_case_, __LINE__));                                            \
    static void CATCH_SHIM_CAT(catch_shim_case_, __LINE__)()

#define REQUIRE(expr)                                                                            \
    do {                                                                                         \
        if (!(expr)) {                                                                           \
            std::printf("  REQUIRE failed: %s (%s:%d)\n", #expr, __FILE__, __LINE__);            \
            throw catch_shim::RequireFailed{};                                                   \
        }                                                                                        \
    } while (0)

int main() {
    int failed_cases = 0;
    for (auto& c: catch_shim::registry()) {
        bool ok = true;
        try {
            c.fn();
        } catch (const catch_shim::RequireFailed&) {
            ok = false;
        } catch (const std::exception& e) {
            std::printf("  exception: %s\n", e.what());
            ok = false;
        }
        std::printf("[%s] %s\n", ok ? "PASS" : "FAIL", c.name);
        failed_cases += ok ? 0 : 1;
    }
    std::printf("%d failures / %zu cases\n", failed_cases, catch_shim::registry().size());
    return failed_cases == 0 ? 0 : 1;
}
