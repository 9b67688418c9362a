// raven-b200: minimal stand-in for cereal (v1.3.0 in the reference,
// Raven.deps.cmake:21-26). Only what RavenLib's public headers need to
// compile: `friend cereal::access` (raven/pile.h:114) and CEREAL_NVP.
#ifndef CEREAL_ACCESS_HPP_
#define CEREAL_ACCESS_HPP_
namespace cereal {
class access {
 public:
  template <class T>
  static T* construct() { return new T(); }
  template <class Archive, class T>
  static void member_serialize(Archive& ar, T& t) { t.serialize(ar); }
};
}  // namespace cereal
#endif  // CEREAL_ACCESS_HPP_
