// oracle/ref_shim/boost/optional.hpp — TEST INFRASTRUCTURE ONLY: the slice of
// boost::optional that limbo/opt/optimizer.hpp uses, on top of std::optional.
#pragma once
#include <optional>
namespace boost {
template <typename T>
struct optional : std::optional<T> {
    using std::optional<T>::optional;
    optional() = default;
    optional(const T& v) : std::optional<T>(v) {}
    const T& get() const { return this->value(); }
    T& get() { return this->value(); }
    bool is_initialized() const { return this->has_value(); }
};
} // namespace boost
