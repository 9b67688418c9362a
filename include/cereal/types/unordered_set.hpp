// raven-b200: cereal/types/unordered_set.hpp - the container support lives in the
// archive classes of our own small cereal (cereal/archives/*.hpp).
#ifndef CEREAL_TYPES_UNORDERED_SET_HPP_
#define CEREAL_TYPES_UNORDERED_SET_HPP_
#include "cereal/cereal.hpp"
#endif
