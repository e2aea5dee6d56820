// Tuple typedefs of the reference's defs.h:4-5.
#ifndef _DEFS_H_
#define _DEFS_H_
#include <tuple>
#include <vector>
typedef std::tuple<int, int, float> UIRating;
typedef std::vector<UIRating> UIRatings;
#endif
