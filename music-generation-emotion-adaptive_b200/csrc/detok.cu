// Device-side detokenisation of generated token ids into note events: the loop of reference api_cache.py:208-221
// (per token: "[INSTRUMENT] <name>" opens a NEW instrument; a whole-line "[NOTE] [PITCH:..] [START:..] [END:..] [DURATION:..]" token
// -- the note_re of api_cache.py:157 -- appends a note to the most recent instrument, and is dropped while there is none) as a
// table gather.  The regex, the float() parses and the pretty_midi name look-ups run ONCE per vocabulary entry on the host
// (vocab.py:note_table); here a token id indexes a 16-byte record and one warp per sequence compacts the events in token order.
#include "detok.cuh"

#include "mg_engine.h"

namespace mg {

namespace {

constexpr uint32_t kFull = 0xffffffffu;

// One warp per sequence; 32 tokens per round; running counts carry the "current instrument" across rounds.
__global__ void __launch_bounds__(128)
detok_kernel(const int32_t* __restrict__ out_ids, const int32_t* __restrict__ out_len, int out_stride, const int4* __restrict__ table,
             int V, int B, DetokOut o) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int n = out_len[b];
  const int32_t* ids = out_ids + static_cast<size_t>(b) * out_stride;
  int n_inst = 0, n_note = 0;                                          // warp-uniform running counts
  for (int t0 = 0; t0 < n; t0 += 32) {
    const int t = t0 + lane;
    int4 rec = make_int4(DETOK_OTHER, 0, 0, 0);
    if (t < n) {
      const int id = ids[t];
      if (id >= 0 && id < V) rec = table[id];
    }
    const bool is_inst = rec.x == DETOK_INSTRUMENT, is_note = rec.x == DETOK_NOTE;
    const uint32_t mi = __ballot_sync(kFull, is_inst), below = (1u << lane) - 1u;
    const int inst_before = n_inst + __popc(mi & below);               // instruments opened before this token
    if (is_inst) {
      const int slot = inst_before;
      if (slot < o.max_inst) o.inst_program[static_cast<size_t>(b) * o.max_inst + slot] = rec.y;
      if (slot < o.max_inst) o.inst_token[static_cast<size_t>(b) * o.max_inst + slot] = ids[t];
    }
    const bool keep = is_note && inst_before > 0;                       // "and current_inst" (api_cache.py:215)
    const uint32_t mk = __ballot_sync(kFull, keep);
    if (keep) {
      const int slot = n_note + __popc(mk & below);
      if (slot < o.max_notes) {
        const size_t at = static_cast<size_t>(b) * o.max_notes + slot;
        o.note_inst[at] = inst_before - 1;
        o.note_pitch[at] = rec.y;
        o.note_start[at] = __int_as_float(rec.z);
        o.note_end[at] = __int_as_float(rec.w);
      }
    }
    n_inst += __popc(mi);
    n_note += __popc(mk);
  }
  if (lane == 0) { o.n_inst[b] = n_inst; o.n_notes[b] = n_note; }
}

}  // namespace

int launch_detok(cudaStream_t s, const int32_t* out_ids, const int32_t* out_len, int out_stride, const int4* table, int V, int B,
                 const DetokOut& o) {
  detok_kernel<<<(B + 3) / 4, 128, 0, s>>>(out_ids, out_len, out_stride, table, V, B, o);
  MG_LAUNCH_CHECK();
  return MG_OK;
}

}  // namespace mg
