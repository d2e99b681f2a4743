// TEST INFRASTRUCTURE ONLY.  pybind11 binding of the REFERENCE's own pcl_loss CPU op, compiled from the reference's sources where
// they lie (projects/WSL/wsl/layers/csrc/pcl_loss/pcl_loss_cpu.cpp + pcl_loss.h; the reference binds the same two functions in
// projects/WSL/wsl/layers/csrc/vision.cpp:11-12).  Built by oracle/build_ref.py into oracle/_ref/ (git-ignored); used to run the
// unmodified PCL head for the golden vectors and to pin oracle/pcl_oracle.py.
#include <torch/extension.h>

#include "pcl_loss/pcl_loss.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("pcl_loss_forward", &wsl::pcl_loss_forward, "pcl_loss_forward");
  m.def("pcl_loss_backward", &wsl::pcl_loss_backward, "pcl_loss_backward");
}
