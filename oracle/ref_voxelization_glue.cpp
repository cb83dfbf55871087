// TEST INFRASTRUCTURE (oracle/): link glue for building the reference's own CPU voxelizer as oracle/_ref.
//
// /root/reference/orv/ops/voxelize/voxelization_cpu.cpp is compiled where it lies (oracle/build_ref.py); it
// declares the two device dispatchers below (voxelization_cpu.cpp:190-201) but only voxelization.cpp, which also
// needs the CUDA launchers, defines them, so the reference's CPU-only JIT build (voxelization.py:27-38) fails to
// load with an undefined symbol.  This file supplies the two missing definitions and nothing else: each forwards to
// the reference's CPU implementation (voxelization_cpu.cpp:110-131, :133-173), which is what the reference's
// registry would select for CPU tensors (REGISTER_DEVICE_IMPL at voxelization_cpu.cpp:203-206).
#include <torch/extension.h>

#include <vector>

int hard_voxelize_forward_cpu(const at::Tensor& points, at::Tensor& voxels, at::Tensor& coors,
                              at::Tensor& num_points_per_voxel, const std::vector<float> voxel_size,
                              const std::vector<float> coors_range, const int max_points, const int max_voxels,
                              const int NDim);
void dynamic_voxelize_forward_cpu(const at::Tensor& points, at::Tensor& coors, const std::vector<float> voxel_size,
                                  const std::vector<float> coors_range, const int NDim);

int hard_voxelize_forward_impl(const at::Tensor& points, at::Tensor& voxels, at::Tensor& coors,
                               at::Tensor& num_points_per_voxel, const std::vector<float> voxel_size,
                               const std::vector<float> coors_range, const int max_points, const int max_voxels,
                               const int NDim) {
  return hard_voxelize_forward_cpu(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range, max_points,
                                   max_voxels, NDim);
}

void dynamic_voxelize_forward_impl(const at::Tensor& points, at::Tensor& coors, const std::vector<float> voxel_size,
                                   const std::vector<float> coors_range, const int NDim) {
  dynamic_voxelize_forward_cpu(points, coors, voxel_size, coors_range, NDim);
}
