// crop.cu -- the viewpoint crop of misc.seprate_point_cloud (utils/misc.py:205-256) as ONE launch.
//
// The reference walks the batch in a Python loop and, per cloud: distance of every point to a random viewpoint
// (torch.norm) -> argsort -> the num_crop nearest points become the "crop", the rest the "input" -> fps() on each side.
// Per training step that is B x (norm + argsort + 2-3 gathers + 2 batch-1 FPS launches)
// (tools/runner_module.py:131, tools/runner_pretask.py:179, tools/runner_unify_seg.py:212).  Here: one CTA per cloud,
//   * the cloud staged once in shared memory (TMA bulk copy);
//   * one 64-bit key per point, (bits of the Euclidean distance << 32) | point index -- distances are non-negative, so
//     the unsigned key order is (distance, index): exactly a STABLE ascending sort on the distance;
//   * a bitonic sort of the keys: 8 keys per thread in registers, distances up to 128 by warp shuffles, only the
//     distances >= 256 through shared memory (8192 keys: 64 KB, 21 barrier-separated passes for the 91 stages);
//   * the split at num_crop and both gathers straight into the two FPS input buffers (or, padding_zeros, the cloud
//     with its cropped rows multiplied by zero, as the reference does).
// The two FPS calls of the whole batch follow as two launches (the larger side on clusters of CTAs, fps.cu).
// Distance: sqrt_rn(fma(dz,dz,fma(dy,dy,dx*dx))) with d* = viewpoint - point -- torch.norm's sum of squares in x,y,z
// order followed by the square root; the ordering is what matters, and the golden fixtures made by the reference's
// own function (tests/golden/golden_seprate.npz) pin it.
#include "bitonic.cuh"

namespace upp {

constexpr int kCropMaxPoints = 8192;
constexpr int kCropThreads = 1024;

__global__ void __launch_bounds__(kCropThreads, 1)
    crop_split_kernel(const float* __restrict__ xyz, const float* __restrict__ centers, int n, int npow2, int num_crop,
                      int padding_zeros, float* __restrict__ crop_out, float* __restrict__ input_out,
                      int32_t* __restrict__ order_out) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  unsigned long long* s_key = reinterpret_cast<unsigned long long*>(s_raw);           // npow2 keys
  float* s_xyz = reinterpret_cast<float*>(s_raw + static_cast<size_t>(npow2) * 8);    // 3*n floats
  __shared__ __align__(8) uint64_t s_bar;
  const int t = threadIdx.x;
  const int b = blockIdx.x;
  const float* p = xyz + static_cast<size_t>(b) * n * 3;
  if (t == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  unsigned parity = 0;
  stage_points(s_xyz, p, n, &s_bar, parity);
  const float cx = __ldg(centers + 3 * b), cy = __ldg(centers + 3 * b + 1), cz = __ldg(centers + 3 * b + 2);
  for (int i = t; i < npow2; i += kCropThreads) {
    unsigned long long key = ~0ull;  // padding sorts last
    if (i < n) {
      const float dx = __fsub_rn(cx, s_xyz[3 * i]), dy = __fsub_rn(cy, s_xyz[3 * i + 1]), dz = __fsub_rn(cz, s_xyz[3 * i + 2]);
      const float d = __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
      key = (static_cast<unsigned long long>(__float_as_uint(d)) << 32) | static_cast<unsigned>(i);
    }
    s_key[i] = key;
  }
  __syncthreads();
  // bitonic sort, ascending (bitonic.cuh: 8 keys per thread in registers, shuffles up to distance 128, shared memory beyond)
  bitonic_sort_cta<unsigned long long, 8>(s_key, npow2);
  // split + gathers
  const int n_in = n - num_crop;
  float* crop_b = crop_out + static_cast<size_t>(b) * num_crop * 3;
  float* in_b = input_out + static_cast<size_t>(b) * (padding_zeros ? n : n_in) * 3;
  for (int r = t; r < n; r += kCropThreads) {
    const int src = static_cast<int>(static_cast<unsigned>(s_key[r]));
    const float x = s_xyz[3 * src], y = s_xyz[3 * src + 1], z = s_xyz[3 * src + 2];
    if (order_out) order_out[static_cast<size_t>(b) * n + r] = src;
    if (r < num_crop) {
      crop_b[3 * r] = x; crop_b[3 * r + 1] = y; crop_b[3 * r + 2] = z;
      if (padding_zeros) {  // input_data[idx[:num_crop]] = input_data[idx[:num_crop]] * 0 (utils/misc.py:236)
        in_b[3 * src] = __fmul_rn(x, 0.f); in_b[3 * src + 1] = __fmul_rn(y, 0.f); in_b[3 * src + 2] = __fmul_rn(z, 0.f);
      }
    } else if (padding_zeros) {
      in_b[3 * src] = x; in_b[3 * src + 1] = y; in_b[3 * src + 2] = z;
    } else {
      const int q = r - num_crop;
      in_b[3 * q] = x; in_b[3 * q + 1] = y; in_b[3 * q + 2] = z;
    }
  }
}

int crop_split_launch(const float* xyz, const float* centers, int B, int n, int num_crop, int padding_zeros,
                      float* crop_out, float* input_out, int32_t* order_out, cudaStream_t st) {
  if (n > kCropMaxPoints) return UPP_ERR_UNSUPPORTED;
  int npow2 = 256;  // a warp sorts blocks of 256 keys in registers
  while (npow2 < n) npow2 <<= 1;
  const size_t smem = static_cast<size_t>(npow2) * 8 + static_cast<size_t>(n) * 12 + 16;
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(crop_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  crop_split_kernel<<<B, kCropThreads, smem, st>>>(xyz, centers, n, npow2, num_crop, padding_zeros, crop_out, input_out,
                                                   order_out);
  count_launch();
  return launch_status();
}

}  // namespace upp
