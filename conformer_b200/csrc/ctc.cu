// Greedy CTC decode on the device (processing/processor.py:301-328 of the reference, id level):
//   ids[t] = argmax_v logits[b, t, v]  (first maximum, like torch.argmax);
//   frames whose id is the pad (blank) or the unk id are skipped WITHOUT resetting the previous id;
//   an id is emitted when it differs from the previously emitted one.
// (So "a blank a" collapses to a single "a": non-standard CTC, kept because the reference does it.)
// Like the reference, every row is decoded over the full padded length T.
//
// One CTA per utterance.  Phase 1: one warp per frame, lanes stride over the vocabulary (coalesced), shuffle reduction of
// (value, index) with the lowest index winning ties.  Phase 2: warp 0 walks the T ids in chunks of 32: ballot of the
// non-skipped frames, each lane fetches the id of the nearest non-skipped frame before it (shuffle inside the chunk, a
// warp-uniform carry across chunks), keeps its frame if the ids differ, and compacts in place with a popcount prefix.
// HBM-bound: B*T*V*4 bytes read once.
#include "ops.h"

namespace cfm {
namespace {

// should candidate (xb, ib) replace the running best (xa, ia)?  Larger value wins, NaN counts as the largest
// (torch.argmax), equal values (or two NaNs) are decided by the lower index; an index of INT_MAX means "nothing yet".
__device__ __forceinline__ bool argmax_replaces(float xa, int ia, float xb, int ib) {
  if (ib == 0x7fffffff) return false;
  if (ia == 0x7fffffff) return true;
  const bool a_nan = xa != xa, b_nan = xb != xb;
  if (a_nan || b_nan) return (a_nan && b_nan) ? ib < ia : b_nan;
  return xb > xa || (xb == xa && ib < ia);
}

__global__ void __launch_bounds__(256) ctc_greedy_kernel(const float* __restrict__ logits, int T, int V, int pad_id,
                                                         int unk_id, int32_t* __restrict__ tokens,
                                                         int32_t* __restrict__ counts) {
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float* lg = logits + static_cast<int64_t>(b) * T * V;
  int32_t* out = tokens + static_cast<int64_t>(b) * T;
  // ---- phase 1: argmax of every frame
  for (int t = warp; t < T; t += nwarps) {
    const float* row = lg + static_cast<int64_t>(t) * V;
    float best = 0.f;
    int bi = 0x7fffffff;
    for (int v = lane; v < V; v += 32) {
      const float x = row[v];
      if (argmax_replaces(best, bi, x, v)) { best = x; bi = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (argmax_replaces(best, bi, ob, oi)) { best = ob; bi = oi; }
    }
    if (lane == 0) out[t] = bi;
  }
  __syncthreads();
  if (warp != 0) return;
  // ---- phase 2: skip pad / unk, collapse repeats, compact in place
  int count = 0;
  int prev = -1;                      // id of the last non-skipped frame seen so far (-1: none yet); warp-uniform
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const int id = t < T ? out[t] : pad_id;
    const bool valid = t < T && id != pad_id && id != unk_id;
    const uint32_t vm = __ballot_sync(0xffffffffu, valid);
    const uint32_t below = vm & ((1u << lane) - 1u);
    const int src = below ? 31 - __clz(below) : 0;
    const int pid_in = __shfl_sync(0xffffffffu, id, src);
    const int pid = below ? pid_in : prev;
    const bool keep = valid && id != pid;
    const uint32_t km = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) out[count + __popc(km & ((1u << lane) - 1u))] = id;
    count += __popc(km);
    if (vm) prev = __shfl_sync(0xffffffffu, id, 31 - __clz(vm));
  }
  __syncwarp();
  for (int t = count + lane; t < T; t += 32) out[t] = -1;
  if (lane == 0) counts[b] = count;
}

}  // namespace

int ctc_greedy_decode(const float* logits, int B, int T, int V, int pad_id, int unk_id, int32_t* tokens, int32_t* counts,
                      cudaStream_t stream) {
  CFM_REQUIRE(B > 0 && T > 0 && V > 0, "bad CTC decode shape");
  ctc_greedy_kernel<<<B, 256, 0, stream>>>(logits, T, V, pad_id, unk_id, tokens, counts);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfm
