// Test double of the reference's body.h (layout only): position, velocity, mass.
#ifndef BODY_H
#define BODY_H
#include "vector.h"
template <int D>
struct Body {
    Vector<D> position;
    Vector<D> velocity;
    double mass = 0.0;
};
#endif
