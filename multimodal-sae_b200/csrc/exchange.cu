// Peer-memory all-gather of small per-token lists for the feature-sharded scan.
//
// Under feature sharding every token chunk needs two exchanges of [Tc, m] fp32 lists per rank (lower bounds, then exact
// local TopK values).  NCCL's all-gather kernels cannot co-reside with the persistent GEMM CTAs of the next chunk
// (registers), so on the second stream they wait for a GEMM launch boundary and then displace GEMM CTAs.  This kernel
// is small enough to run beside them: every rank PUSHES its slab straight into the gathered buffer of every peer
// (plain 16-byte stores to mapped peer pointers over NVLink, or ONE multimem.st per vector when the buffer has an
// NVSwitch multicast mapping), then the last CTA to finish publishes a per-(channel, source rank) sequence number on
// every peer with a system-scope release store and spins (acquire loads) until all peers' numbers have arrived.  When
// the kernel ends, the local gathered buffer [R][Tc][m] is complete and the next kernel in the stream may read it.
//
// The buffers are symmetric allocations (same offsets on every rank; PyTorch's symmetric-memory rendezvous provides
// the peer / multicast mappings -- plumbing only).  Sequence numbers only grow, so the flags are never reset.
// Re-use of a gathered region is safe without an extra barrier because the two exchanges of a chunk alternate: a rank
// can only push exchange e of chunk c+1 after it has seen every peer's push of the OTHER exchange, which those peers
// issue (stream order) after their last read of region e for chunk c.
#include "kernels_exchange.cuh"

namespace saeb {

int push_gather_launch(const void* src, size_t bytes, void* const* peer_bases_dev, int R, int self_rank,
                       size_t region_offset, void* multicast_base, size_t flags_offset, int channel, unsigned int seq,
                       int* counter, cudaStream_t stream) {
  SAEB_REQUIRE(R >= 1 && R <= PUSH_MAX_RANKS && self_rank >= 0 && self_rank < R, "push_gather: bad rank %d of %d",
               self_rank, R);
  SAEB_REQUIRE(bytes % 16 == 0 && region_offset % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0,
               "push_gather: slab size, region offset and source must be 16-byte aligned");
  SAEB_REQUIRE(flags_offset % 4 == 0 && channel >= 0, "push_gather: bad flags offset / channel");
  const size_t n_vec = bytes / 16;
  const size_t slab_off = region_offset + (size_t)self_rank * bytes;
  uint4* mc = multicast_base
                  ? reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(multicast_base) + slab_off)
                  : nullptr;
  size_t blocks = (n_vec + PUSH_THREADS * 4 - 1) / (PUSH_THREADS * 4);
  if (blocks < 1) blocks = 1;
  if (blocks > 64) blocks = 64;
  SAEB_CARVEOUT(push_gather_kernel);
  push_gather_kernel<<<(unsigned)blocks, PUSH_THREADS, 0, stream>>>(reinterpret_cast<const uint4*>(src), n_vec,
                                                                    peer_bases_dev, R, self_rank, slab_off, mc,
                                                                    flags_offset, channel, seq, counter);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace saeb
