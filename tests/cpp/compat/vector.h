// Test double of the reference's vector.h: ONLY the layout and accessors the adapter needs
// (std::array<double,D> components, operator[]).  Used when /root/reference is absent (GPU box);
// where the reference is present the adapter is compiled against its real headers instead.
#ifndef VECTOR_H
#define VECTOR_H
#include <array>
template <int D>
class Vector {
public:
    std::array<double, D> components{};
    double& operator[](int i) { return components[i]; }
    const double& operator[](int i) const { return components[i]; }
};
#endif
