// Test infrastructure only (oracle/): stand-in for re2 (reference CMakeLists.txt:30-35).
// include/statement.h:118-161 uses RE2 for LIKE filters, which run in the harness before
// execute(); nothing on the join path calls it, the header just has to compile.
#pragma once
#include <regex>
#include <string>
#include <string_view>

class RE2 {
public:
    struct Options {};

    RE2(const std::string& pattern, const Options&) {
        try {
            re_ = std::regex(pattern);
            ok_ = true;
        } catch (...) {
            ok_ = false;
        }
    }

    bool ok() const { return ok_; }

    static bool FullMatch(std::string_view s, const RE2& r) {
        return r.ok_ && std::regex_match(s.begin(), s.end(), r.re_);
    }

private:
    std::regex re_;
    bool       ok_ = false;
};
