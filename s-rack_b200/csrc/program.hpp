// The flat device program a planned patch compiles to: one instruction per module
// (per channel for Output) in plan order, plus ring loads/stores for wires whose
// source runs later in the plan than their reader (the reference's one-block
// feedback latency, src/synth.rs:168-192 + :32).  Shared by host and device code.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace srk {

enum Op : uint8_t {
  OP_END = 0,
  OP_RING_LOAD,   // out[0] <- ring[aux] (samples delayed by buffer_size)
  OP_RING_STORE,  // ring[aux] <- in[0]
  OP_OSC,
  OP_NOISE,
  OP_MOOG,
  OP_ADSR,
  OP_VCA,
  OP_MIXER,
  OP_MATH,
  OP_OUTPUT,
};

// Instr::flags
enum : uint8_t {
  F_MATH_ADD = 0, F_MATH_SUB = 1, F_MATH_MUL = 2, F_MATH_NONLIN = 3,
  F_OUT_SAME_AS_PREV = 1,  // this channel reads the same wire as the previous channel's instr
};

// ADSR mode encoding in the state word (adsr.rs:26-33 order)
enum : uint32_t { ADSR_ATTACK = 0, ADSR_DECAY = 1, ADSR_SUSTAIN = 2, ADSR_RELEASE = 3, ADSR_NONE = 4 };

struct alignas(16) Instr {
  uint8_t op;
  uint8_t flags;
  int16_t in[4];    // wire slot per input, -1 = not connected (None)
  int16_t out[3];   // wire slot per output port, -1 = nobody reads it (not materialised)
  uint16_t state;   // first per-voice state word
  uint16_t param;   // first per-voice parameter word
  uint16_t aux;     // ring id / module index (noise key) / channel (output)
  uint16_t pad;
  float imm;        // oscillator / ADSR sample rate
  uint32_t pad2;
};
static_assert(sizeof(Instr) == 32, "Instr must stay 32 bytes (staged to shared memory as uint4 pairs)");

// Per-voice state words (u32 slots, SoA [word][voice] in HBM)
//   OSC   : pos (f64, 2 words), sync_last                         oscillator.rs:21,23
//   NOISE : sample counter (u64, 2 words)
//   MOOG  : f, p, q, b[0..4], freq, res                            filter.rs:48-56
//   ADSR  : phase, r_val, from_a_val, mode | gate_last << 8        adsr.rs:14-21
constexpr int kStateOsc = 3, kStateNoise = 2, kStateMoog = 10, kStateAdsr = 4;
// Per-voice parameter words (SoA [word][voice] in HBM)
//   OSC   : val, delta (f64, 2 words; host-computed 440*2^val/sr, used when CV is None), antialiasing (0/1)
//   MOOG  : freq, res, exp_amt        ADSR : a_sec, d_sec, s_val, r_sec
//   MIXER : gain[0..3]                MATH : constant          VCA : negative (0/1)
constexpr int kParamOsc = 4, kParamMoog = 3, kParamAdsr = 4, kParamMixer = 4, kParamMath = 1, kParamVca = 1;

// Where a parameter word comes from when the table is (re)built for a voice range.
struct ParamSource {
  int module;       // index into the patch's module list at plan time
  int pid;          // srk_param id; -1 => derived oscillator delta (lo word), -2 => (hi word)
};

struct Program {
  std::vector<Instr> code;        // terminated by OP_END
  std::vector<uint32_t> state_init;  // one initial value per state word
  std::vector<ParamSource> param_src;  // one per parameter word
  uint32_t n_wires = 0;           // physical wire slots
  uint32_t n_rings = 0;
  uint32_t channels = 0;
  uint32_t ring_len = 0;          // buffer_size
};

}  // namespace srk
