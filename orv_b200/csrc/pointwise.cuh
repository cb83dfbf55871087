// Internal launch entry points of the kernels in pointwise.cu / gemm.cu / attention.cu, shared with forward.cu.
#pragma once
#include "common.cuh"

namespace orvb {

struct SkinnyJob {
  const bf16* w;  // [n, k]
  const bf16* b;  // [n] or null
  float* y;       // [rows, n]
};

// one AdaLN site for ab_combine: LayerNorm affine (bf16 [dim] or null), its fp32 modulation table and the bf16
// output table [groups][4*dim]
struct AbSite {
  const bf16* ln_w;
  const bf16* ln_b;
  const float* mod;
  int mod_ld, text_off, video_off;
  bf16* ab;
};
int ab_combine_launch(const AbSite* sites_dev, int num_sites, int groups, int dim, cudaStream_t stream);

int ln_modulate_launch(const orvb_ln_args* a, cudaStream_t stream);
int skinny_linear_launch(const float* x, const SkinnyJob& job, const SkinnyJob* jobs_dev, int num_jobs, int rows,
                         int n, int k, int act, cudaStream_t stream);
int patchify_launch(const void* x, void* out, int B, int F, int C, int H, int W, int p, int patch_t,
                    cudaStream_t stream);
int unpatchify_launch(const void* y, void* out, int B, int F, int C, int H, int W, int p, int patch_t,
                      cudaStream_t stream);
int sampler_step_launch(const orvb_sampler_step_args* a, cudaStream_t stream);
int attention_launch(const void* qkv, void* out, int batch, int seq_len, int heads, float scale, int q_row0,
                     int q_rows, cudaStream_t stream, int out_f32 = 0);
// (b v)(text | f s) rows -> (b f)(v text | v s) rows, `width` bf16 per row (MVBlock rearranges, :328-331)
int mv_gather_launch(const void* src, void* dst, int clips, int views, int frames, int text_len, int tokens,
                     int width, cudaStream_t stream);

}  // namespace orvb
