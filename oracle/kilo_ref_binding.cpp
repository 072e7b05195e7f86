// Binding of the REFERENCE's own KiloNeRF kernels for use as a GPU-side checker (test infrastructure).
// The kernels are compiled from the sources where they lie (/root/reference/cuda/{generate_inputs,
// network_eval,integrate,utils}.cu) by oracle/build_kilo_ref.py; nothing of them is copied into this repo.
// Only the four functions on the a9 path are exposed (the reference's own pybind.cu also pulls in the
// MAGMA / GL parts, which cannot be built here).  Declarations come from the reference's headers.
#include <torch/extension.h>

#include "generate_inputs.cuh"
#include "integrate.cuh"
#include "network_eval.cuh"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("get_rays_d", &get_rays_d);
  m.def("generate_query_indices_on_ray", &generate_query_indices_on_ray);
  m.def("network_eval_query_index", &network_eval_query_index);
  m.def("integrate", &integrate);
  m.def("replace_transparency_by_background_color", &replace_transparency_by_background_color);
}
