// row_compact.cuh -- the cold-row delta stream: what the copy engine moves instead of full rows.
//
// Reference behaviour behind it: on a rejected proposal ChainArray::append re-pushes a COPY of the chain's last state
// (src/infer/chainarray.cpp:102-106) with accepted = false, and a swap only rewrites the cold row when the pair
// (0, 1) exchanges (chainarray.cpp:162-189).  So the row a cold chain emits in a round equals its previous row in
// x, energy, sigma and beta about three times in four; only the two flags differ.  k_compact_rows compares every
// row of a launch with the same stack's previous row, bit for bit, and emits
//     flags[round][stack]    bit 0 accepted, bits 1-2 swapType, bit 3 changed
//     payload[slot][d + 3]   x, energy, sigma, beta of the changed rows
//     base[round][chunk]     payload slot of the first changed row of a chunk of 256 stacks (slots inside a chunk
//                            follow the stack order, so the host finds them by counting flags)
// into one contiguous buffer (engine.cu: SegLayout) that a single cudaMemcpyAsync carries to the host.  The host
// rebuilds the exact (d + 5)-double rows (engine.cu: materialise); nothing is lost or approximated.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace slgpu_rows {

constexpr int kChunk = 256;  // stacks per block = per base entry

// grid (n_chunks, n_rounds), 256 threads.  rows: this launch's full rows [n_rounds][S][w] (w = d + 5);
// prev_in: the previous launch's last round [S][w] (ignored when key != 0: every row of round 0 is then "changed");
// prev_out: receives this launch's last round (a different buffer from prev_in: the blocks of round 0 still read that).
// out: the compact segment; its count word (out + 8) must be zero on entry.
__global__ void __launch_bounds__(kChunk) k_compact_rows(const double* __restrict__ rows, const double* __restrict__ prev_in,
                                                         double* __restrict__ prev_out, uint32_t S, uint32_t w, uint32_t n_rounds,
                                                         int key, unsigned char* out, uint32_t n_chunks, uint32_t off_flags,
                                                         uint32_t off_payload) {
  const uint32_t chunk = blockIdx.x, round = blockIdx.y, tid = threadIdx.x;
  const uint32_t s0 = chunk * kChunk;
  const uint32_t nrow = (S - s0 < (uint32_t)kChunk) ? S - s0 : (uint32_t)kChunk;
  const uint32_t wp = w - 2;  // payload doubles per row
  const double* cur = rows + ((size_t)round * S + s0) * w;
  const double* prev = round ? cur - (size_t)S * w : prev_in + (size_t)s0 * w;
  const bool all = (round == 0 && key != 0);
  const bool last_round = (round + 1 == n_rounds);
  __shared__ int s_ch[kChunk];
  __shared__ uint32_t s_slot[kChunk];
  __shared__ uint32_t s_warp[kChunk / 32];
  __shared__ uint32_t s_base;
  s_ch[tid] = (all && tid < nrow) ? 1 : 0;
  __syncthreads();
  const uint32_t n = nrow * w;
  const uint32_t dr = (uint32_t)kChunk / w, dc = (uint32_t)kChunk - dr * w;  // a stride of 256 doubles in (row, column) steps
  {
    uint32_t r = tid / w, c = tid - r * w;
    for (uint32_t i = tid; i < n; i += kChunk) {  // consecutive threads, consecutive doubles
      const double v = cur[i];
      if (!all && c < wp && __double_as_longlong(v) != __double_as_longlong(prev[i])) s_ch[r] = 1;  // same value from every writer
      if (last_round) prev_out[(size_t)s0 * w + i] = v;
      r += dr;
      c += dc;
      if (c >= w) { c -= w; r += 1; }
    }
  }
  __syncthreads();
  // exclusive scan of the changed flags over the chunk: slots follow the stack order
  const int ch = s_ch[tid];
  const uint32_t lane = tid & 31, wid = tid >> 5;
  const uint32_t bal = __ballot_sync(0xffffffffu, ch != 0);
  const uint32_t before = __popc(bal & ((1u << lane) - 1u));
  if (lane == 0) s_warp[wid] = __popc(bal);
  __syncthreads();
  uint32_t off = 0, total = 0;
#pragma unroll
  for (int i = 0; i < kChunk / 32; ++i) {
    if ((uint32_t)i < wid) off += s_warp[i];
    total += s_warp[i];
  }
  s_slot[tid] = off + before;
  if (tid == 0) {
    const uint32_t b = total ? atomicAdd((unsigned int*)(out + 8), total) : 0u;
    s_base = b;
    ((uint32_t*)(out + 16))[(size_t)round * n_chunks + chunk] = b;
    if (chunk == 0 && round == 0) {
      ((uint32_t*)out)[0] = n_rounds;
      ((uint32_t*)out)[1] = S;
      ((uint32_t*)out)[3] = (uint32_t)(key != 0);
    }
  }
  if (tid < nrow) {
    const int acc = (int)cur[(size_t)tid * w + wp], sw = (int)cur[(size_t)tid * w + wp + 1];
    (out + off_flags)[(size_t)round * S + s0 + tid] = (unsigned char)((acc & 1) | ((sw & 3) << 1) | (ch ? 8 : 0));
  }
  __syncthreads();
  if (total == 0) return;
  double* pay = (double*)(out + off_payload) + (size_t)s_base * wp;
  uint32_t r = tid / w, c = tid - r * w;
  for (uint32_t i = tid; i < n; i += kChunk) {
    if (c < wp && s_ch[r]) pay[(size_t)s_slot[r] * wp + c] = cur[i];
    r += dr;
    c += dc;
    if (c >= w) { c -= w; r += 1; }
  }
}

}  // namespace slgpu_rows
