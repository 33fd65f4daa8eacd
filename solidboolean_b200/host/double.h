// Epsilon helpers with the reference's semantics (reference src/double.h:31-39).
#ifndef SB_HOST_DOUBLE_H
#define SB_HOST_DOUBLE_H
#include <cmath>
#include <limits>

namespace Double
{
inline bool isZero(double value) { return std::fabs(value) <= std::numeric_limits<double>::epsilon(); }
inline bool isEqual(double a, double b) { return isZero(a - b); }
}

#endif
