// TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory).  Thread-block clusters are not
// emulated: the cluster kernels compile against these stubs but are never launched -- the emulator
// runs with RSG_SCB_NO_CLUSTER=1, i.e. the one-CTA-per-sub-problem 4-colour SOR kernel, to which the
// cluster kernels are bit-identical on the device (tests/test_scb_parity_gpu.py).
#pragma once
#include <cstdlib>
namespace cooperative_groups {
struct cluster_group {
  unsigned num_blocks() const { return 1; }
  unsigned block_rank() const { return 0; }
  void sync() const { abort(); }
  template <class T> T* map_shared_rank(T* p, unsigned) const { abort(); return p; }
};
inline cluster_group this_cluster() { return cluster_group(); }
}  // namespace cooperative_groups
