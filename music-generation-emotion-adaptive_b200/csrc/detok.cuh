// Device-side detokenisation (implementation: detok.cu; reference loop: api_cache.py:157,208-221).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace mg {

enum DetokKind { DETOK_OTHER = 0, DETOK_INSTRUMENT = 1, DETOK_NOTE = 2 };

// Per-token record (int4): kind, program (instrument) or MIDI pitch (note), start and end seconds as float bits.
struct DetokOut {
  int32_t* n_inst;        // [B] instruments opened (may exceed max_inst: the caller sees the overflow)
  int32_t* inst_program;  // [B][max_inst] GM program of the i-th "[INSTRUMENT]" token
  int32_t* inst_token;    // [B][max_inst] its token id (the name stays a host-side look-up)
  int32_t* n_notes;       // [B] notes kept (may exceed max_notes)
  int32_t* note_inst;     // [B][max_notes] index into the instrument list
  int32_t* note_pitch;    // [B][max_notes]
  float* note_start;      // [B][max_notes]
  float* note_end;        // [B][max_notes]
  int32_t max_inst, max_notes;
};

int launch_detok(cudaStream_t s, const int32_t* out_ids, const int32_t* out_len, int out_stride, const int4* table, int V, int B,
                 const DetokOut& o);

}  // namespace mg
