// Test infrastructure only (oracle/): minimal stand-in for range-v3 0.12.0, which the reference
// pins in CMakeLists.txt:51-56 but which cannot be fetched offline.  Maps the handful of range-v3
// names used by /root/reference/src/{execute,build_table}.cpp and include/table.h onto C++23 <ranges>.
#pragma once
#include <ranges>
#include <string>
#include <utility>
#include <vector>

namespace ranges {
namespace views {
using std::views::enumerate;
using std::views::transform;
using std::views::zip;
inline constexpr auto join = [](auto sep) { return std::views::join_with(sep); };
} // namespace views

template <class Container>
struct to_closure {};

template <class Container>
to_closure<Container> to() {
    return {};
}

template <class Range, class Container>
Container operator|(Range&& r, to_closure<Container>) {
    Container c;
    for (auto&& x: r) {
        c.insert(c.end(), std::forward<decltype(x)>(x));
    }
    return c;
}
} // namespace ranges
