// TEST INFRASTRUCTURE ONLY -- C entry points that run the device code of multimodal-sae_b200/csrc/kernels_*.cuh on the
// host through tests/emu/cuda_emu.h.  Built by tests/test_kernel_emu.py with g++ from a copy of the kernel headers in
// which `extern __shared__` / `__shared__` were replaced by `extern` / `static`.
#define SAEB_CPU_EMU 1
#include <cstdarg>

#include "kernels_decode_bwd.cuh"
#include "kernels_exchange.cuh"
#include "kernels_kth.cuh"
#include "kernels_pack.cuh"
#include "kernels_refine.cuh"

namespace saeb {
void set_error(const char*, ...) {}
// dynamic shared memory of the emulated kernels (`extern __shared__ T name[]` in the sources)
alignas(16) float rsm[1 << 16];
alignas(16) float bsm[1 << 12];
alignas(16) float gsm[1 << 16];
}  // namespace saeb

using namespace saeb;

extern "C" {

void emu_kth(int vpl, const float* g, int R, long long T, int m, int kth, float* out) {
  const unsigned blocks = (unsigned)((T + 7) / 8);
  emu::launch({blocks}, {256}, [&] {
    if (vpl == 4) kth_gathered_reg_kernel<4>(g, R, T, m, kth, out);
    else if (vpl == 16) kth_gathered_reg_kernel<16>(g, R, T, m, kth, out);
    else if (vpl == 64) kth_gathered_reg_kernel<64>(g, R, T, m, kth, out);
    else kth_gathered_mem_kernel(g, R, T, m, kth, out);
  });
}

void emu_decode_bwd_acts(const float* g, long long ld_g, const long long* idx, long long T, int k, const float* W,
                         long long d, long long N, float* d_vals, int* err_flag) {
  emu::launch({(unsigned)T}, {(unsigned)DBW_THREADS},
              [&] { decode_bwd_acts_kernel(g, ld_g, idx, k, W, d, N, d_vals, err_flag); });
}

void emu_decode_bwd_weight(const float* g, long long ld_g, const long long* idx, const float* vals, long long T, int k,
                           long long d, long long N, float* dW, int* err_flag) {
  emu::launch({(unsigned)T}, {(unsigned)DBW_THREADS},
              [&] { decode_bwd_weight_kernel(g, ld_g, idx, vals, k, d, N, dW, err_flag); });
}

// mode 3 / 4 weight pack exactly as pack_weights_f16_launch + pack_weights_lo_launch enqueue it
void emu_pack(const float* W, const float* b_enc, const float* b_dec, long long N, long long d, long long d_pad,
              void* hi, void* lo, float* bias, float* wnorm, float* dnorm, float* trailer) {
  std::memset(trailer, 0, 16);
  const unsigned blocks = (unsigned)((N + 7) / 8);
  emu::launch({blocks}, {256}, [&] {
    w_stats_kernel(W, b_enc, b_dec, N, d, bias, wnorm, reinterpret_cast<unsigned int*>(trailer + 2),
                   reinterpret_cast<unsigned int*>(trailer + 1));
  });
  emu::launch({blocks}, {256}, [&] {
    pack_w_f16_kernel(W, N, d, d_pad, reinterpret_cast<unsigned int*>(trailer + 2), reinterpret_cast<__half*>(hi),
                      dnorm, trailer);
  });
  if (lo != nullptr)
    emu::launch({4}, {256}, [&] { pack_w_lo_kernel(W, N, d, d_pad, trailer, reinterpret_cast<__half*>(lo)); });
}

// bf16 activations (raw 16-bit words) -> scaled fp16 plane + row scale + norms
void emu_prep_x_bf16(const void* x, long long T, long long d, long long ld_x, long long d_pad, void* out,
                     float* row_scale, float* xnorm, float* xdnorm) {
  emu::launch({(unsigned)((T + 7) / 8)}, {256}, [&] {
    prep_x_f16_kernel<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(x), T, d, ld_x, d_pad,
                                     reinterpret_cast<__half*>(out), row_scale, xnorm, xdnorm);
  });
}

void emu_candidate_bounds(const float* cand_vals, const long long* cand_idx, long long T, int K2, int k,
                          const float* wnorm, const float* dnorm, const float* xnorm, const float* xdnorm, float c_eps,
                          long long clamp_feature, float* lb_out) {
  emu::launch({(unsigned)T}, {128}, [&] {
    candidate_bounds_kernel(cand_vals, cand_idx, K2, k, wnorm, dnorm, xnorm, xdnorm, c_eps, clamp_feature, lb_out);
  });
}

// the refinement kernel on bf16 activations: lo == nullptr -> exact fp32 re-evaluation (refine_kernel), else the
// residual-plane correction (refine_lo_kernel)
void emu_refine_bf16(const void* x, long long T, long long ld_x, const float* W, long long d, long long N,
                     const float* bias, const float* wnorm, const float* dnorm, const float* trailer,
                     const float* xnorm, const float* xdnorm, float c_eps, const float* cand_vals,
                     const long long* cand_idx, int K2, int k, long long clamp_feature, float clamp_value,
                     float* out_vals, long long* out_idx, int* status, int* flag_rows, const float* ext_lower,
                     const void* lo, long long ld_w, int threads) {
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  emu::launch({(unsigned)T}, {(unsigned)threads}, [&] {
    if (lo == nullptr)
      refine_kernel<__nv_bfloat16>(xb, ld_x, W, d, N, bias, wnorm, dnorm, trailer, xnorm, xdnorm, c_eps, cand_vals,
                                   cand_idx, K2, k, clamp_feature, clamp_value, out_vals, out_idx, status, flag_rows,
                                   ext_lower);
    else
      refine_lo_kernel<__nv_bfloat16>(xb, ld_x, reinterpret_cast<const __half*>(lo), ld_w, d, N, bias, wnorm, dnorm,
                                      trailer, xnorm, xdnorm, c_eps, cand_vals, cand_idx, K2, k, clamp_feature,
                                      clamp_value, out_vals, out_idx, status, flag_rows, ext_lower);
  });
}

// one rank of the peer-memory all-gather, launched like push_gather_launch does (`blocks` CTAs); `peer_bases` holds
// this process' mappings of all R symmetric buffers
void emu_push_gather(const void* src, size_t bytes, void* const* peer_bases, int R, int self, size_t region_offset,
                     size_t flags_offset, int channel, unsigned seq, int* counter, int blocks) {
  const size_t n_vec = bytes / 16;
  const size_t slab_off = region_offset + (size_t)self * bytes;
  emu::launch({(unsigned)blocks}, {(unsigned)PUSH_THREADS}, [&] {
    push_gather_kernel(reinterpret_cast<const uint4*>(src), n_vec, peer_bases, R, self, slab_off, nullptr, flags_offset,
                       channel, seq, counter);
  });
}

}  // extern "C"
