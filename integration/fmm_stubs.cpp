// fmm_stubs.cpp -- ours, not the reference's: definitions for the six FMM members that the reference HEAD
// declares but never defines (fmm.h:98,101: FMMNode<D>::translate_local_to_children / compute_direct_forces;
// fmm_omp.h:36-45: FMM_OMP<D>::{m2l,l2l,l2p,p2p}_phase), so that the suite links at all.  They are
// off the brute-force path and throw if reached (safely_execute logs the exception and skips the row).
// Added to SOURCES by the patched Makefile (integration/build_patched_reference.py).
#include <stdexcept>
#include <vector>

#include "methods.h"

template <int D> void FMMNode<D>::translate_local_to_children(int) {
    throw std::logic_error("FMMNode::translate_local_to_children is not defined in the reference");
}
template <int D>
void FMMNode<D>::compute_direct_forces(std::vector<Vector<D>>&, const std::vector<Body<D>>&) {
    throw std::logic_error("FMMNode::compute_direct_forces is not defined in the reference");
}
template <int D> void FMM_OMP<D>::m2l_phase() { throw std::logic_error("FMM_OMP::m2l_phase is not defined in the reference"); }
template <int D> void FMM_OMP<D>::l2l_phase() { throw std::logic_error("FMM_OMP::l2l_phase is not defined in the reference"); }
template <int D>
void FMM_OMP<D>::l2p_phase(std::vector<Vector<D>>&, const std::vector<Body<D>>&) {
    throw std::logic_error("FMM_OMP::l2p_phase is not defined in the reference");
}
template <int D>
void FMM_OMP<D>::p2p_phase(std::vector<Vector<D>>&, const std::vector<Body<D>>&) {
    throw std::logic_error("FMM_OMP::p2p_phase is not defined in the reference");
}

template void FMMNode<2>::translate_local_to_children(int);
template void FMMNode<3>::translate_local_to_children(int);
template void FMMNode<2>::compute_direct_forces(std::vector<Vector<2>>&, const std::vector<Body<2>>&);
template void FMMNode<3>::compute_direct_forces(std::vector<Vector<3>>&, const std::vector<Body<3>>&);
template void FMM_OMP<2>::m2l_phase();
template void FMM_OMP<3>::m2l_phase();
template void FMM_OMP<2>::l2l_phase();
template void FMM_OMP<3>::l2l_phase();
template void FMM_OMP<2>::l2p_phase(std::vector<Vector<2>>&, const std::vector<Body<2>>&);
template void FMM_OMP<3>::l2p_phase(std::vector<Vector<3>>&, const std::vector<Body<3>>&);
template void FMM_OMP<2>::p2p_phase(std::vector<Vector<2>>&, const std::vector<Body<2>>&);
template void FMM_OMP<3>::p2p_phase(std::vector<Vector<3>>&, const std::vector<Body<3>>&);
