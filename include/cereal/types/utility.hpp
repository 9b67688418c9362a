// raven-b200: cereal/types/utility.hpp - the container support lives in the
// archive classes of our own small cereal (cereal/archives/*.hpp).
#ifndef CEREAL_TYPES_UTILITY_HPP_
#define CEREAL_TYPES_UTILITY_HPP_
#include "cereal/cereal.hpp"
#endif
