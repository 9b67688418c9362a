// raven-b200: own small implementation of the part of `cereal` (v1.3.0 in the
// reference, Raven.deps.cmake:21-26) that RavenLib uses: checkpoints
// (RavenLib/src/binary.cc:23-96: cereal::BinaryOutputArchive / BinaryInputArchive
// over Graph, Pile, Node, Edge, biosoup::NucleicAcid) and the pile dump
// (RavenLib/src/graph_repr.cc:400-416: cereal::JSONOutputArchive + make_nvp).
// Serialisable here: arithmetic types, std::string, std::vector (incl.
// vector<bool>), std::pair, std::unique_ptr, std::unordered_set, name-value
// pairs, and classes with a member `serialize(Archive&)` (through
// cereal::access) or a free `serialize(Archive&, T&)` found by ADL.
// The binary layout follows cereal's (sizes as 64-bit counts, arithmetic values
// raw, contiguous arithmetic vectors as one block, unique_ptr as a validity
// byte + object), so a checkpoint written here reads back here; byte
// compatibility with upstream files is not claimed.
#ifndef CEREAL_CEREAL_HPP_
#define CEREAL_CEREAL_HPP_

#include <cstdint>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_set>
#include <utility>
#include <vector>

#include "cereal/access.hpp"

namespace cereal {

using size_type = std::uint64_t;

struct Exception : public std::runtime_error {
  explicit Exception(const std::string& what) : std::runtime_error(what) {}
};

template <class T>
struct NameValuePair {
  const char* name;
  T& value;
};

template <class T>
inline NameValuePair<T> make_nvp(const char* name, T& value) {
  return {name, value};
}
template <class T>
inline NameValuePair<T> make_nvp(const std::string& name, T& value) {
  return {name.c_str(), value};
}

#ifndef CEREAL_NVP
#define CEREAL_NVP(x) ::cereal::make_nvp(#x, x)
#endif

// Visitor over the fields a class hands to its archive: f(raw references...).
// (raven-b200 extension: how the device-computed pile histograms get into
// raven::Pile, whose only open door is `friend cereal::access` + serialize();
// include/raven_b200/construct_b200.hpp)
template <class T>
inline T& Unwrap(NameValuePair<T> p) { return p.value; }
template <class T>
inline T& Unwrap(T& t) { return t; }
template <class F>
struct FieldVisitor {
  F& f;
  template <class... Ts>
  void operator()(Ts&&... a) { f(Unwrap(a)...); }
};
template <class F>
inline FieldVisitor<F> fields(F& f) { return {f}; }

namespace detail {

template <class A, class T, class = void>
struct has_member_serialize : std::false_type {};
template <class A, class T>
struct has_member_serialize<
    A, T, std::void_t<decltype(access::member_serialize(std::declval<A&>(), std::declval<T&>()))>>
    : std::true_type {};

template <class A, class T, class = void>
struct has_free_serialize : std::false_type {};
template <class A, class T>
struct has_free_serialize<A, T,
                          std::void_t<decltype(serialize(std::declval<A&>(), std::declval<T&>()))>>
    : std::true_type {};

template <class T>
struct is_nvp : std::false_type {};
template <class T>
struct is_nvp<NameValuePair<T>> : std::true_type {};

}  // namespace detail

// CRTP base of the archives: archive(a, b, c...) visits every argument
template <class Derived, bool kLoading>
class ArchiveBase {
 public:
  static constexpr bool is_loading = kLoading;
  static constexpr bool is_saving = !kLoading;

  template <class... Ts>
  Derived& operator()(Ts&&... args) {
    (self().Process(std::forward<Ts>(args)), ...);
    return self();
  }

 protected:
  Derived& self() { return static_cast<Derived&>(*this); }

  // classes: member serialize first, then a free one
  template <class T>
  void Object(T& t) {
    using U = std::remove_const_t<T>;
    U& u = const_cast<U&>(t);
    if constexpr (detail::has_member_serialize<Derived, U>::value) {
      access::member_serialize(self(), u);
    } else if constexpr (detail::has_free_serialize<Derived, U>::value) {
      serialize(self(), u);
    } else {
      static_assert(sizeof(U) == 0, "cereal (raven-b200): type has no serialize function");
    }
  }
};

}  // namespace cereal

#endif  // CEREAL_CEREAL_HPP_
