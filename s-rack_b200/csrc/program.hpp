// The flat device program a planned patch compiles to: one instruction per module
// (one for Output) in plan order, plus ring loads/stores for wires whose source
// runs later in the plan than their reader (the reference's one-block feedback
// latency, src/synth.rs:168-192 + :32).  Shared by host and device code.
//
// Scheduling (v2).  A 32-voice group is rendered by S warps.  Every instruction
// carries the warp that executes it and a pipeline `stage`: at iteration i the
// instruction works on chunk (i - stage) of K samples.  A reader's stage is
// strictly greater than its writer's, so the tile it reads was finished in an
// earlier iteration (one block barrier per iteration); each wire is a ring of
// `mask + 1` tiles indexed by chunk & mask.  With S == 1 all stages are 0, the
// instructions run in plan order and wires share single tiles by liveness.
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
  OP_OUTPUT,      // channels aux .. aux+3 <- in[0..3]: per-voice stems
  OP_MIX,         // channels aux .. aux+3: this voice group's share of the mixdown
  OP_MOOG_COEF,   // out[0..2] <- ladder coefficients (f, p, q) of a filter's CV input in[0]
  OP_GRIDSEQ,     // grid sequencer: in step, sync; out cv, gate, sync; table at aux, n_ch steps
  OP_PATSEQ,      // pattern sequencer: in step, sync; out ports flags..flags+2 of its 9; table at aux
  OP_OSC_DELTA,   // out[0..1] <- lo / hi words of delta = 440 * 2^(cv + val) / sample_rate per sample (f64)
  OP_SAMPLE,      // sample player: in gate, cv; out[0]; descriptor (WaveDesc) at tables[aux]
  OP_OSC_PHASE,   // the oscillator's phase recurrence alone: in[1] sync, in[2..3] delta (n_ch = 1); out[0..1] <- lo / hi
                  // words of the phase each sample is shaped at (f64); owns the oscillator's state
  OP_OSC_SHAPE,   // the stateless rest: in[0..1] phase, in[2..3] delta (n_ch = 1, else the voice's constant); out sine, square, saw
};

// Instr::flags
enum : uint8_t { F_MATH_ADD = 0, F_MATH_SUB = 1, F_MATH_MUL = 2, F_MATH_NONLIN = 3 };
enum : uint8_t { F_MOOG_EXT_COEF = 1 };
// Uniform-only parameters ride in the instruction instead of taking a per-voice word (shared memory is what
// limits the chunk length): OP_OSC bit 3 (bits 0-2 copy index, 4-7 copies), OP_VCA bit 0.
enum : uint8_t { F_OSC_NO_ANTIALIASING = 8, F_VCA_NEGATIVE = 1 };  // OP_MOOG: in[1..3] carry (f, p, q) from an OP_MOOG_COEF instead of the CV

// ADSR mode encoding in the state word (adsr.rs:26-33 order)
enum : uint32_t { ADSR_ATTACK = 0, ADSR_DECAY = 1, ADSR_SUSTAIN = 2, ADSR_RELEASE = 3, ADSR_NONE = 4 };

constexpr int kVoicesPerGroup = 32;  // one lane per voice
constexpr int kMaxWarps = 16;        // warps per group (512 threads x 128 registers)
constexpr int kOutputChannelsPerInstr = 4;

struct alignas(16) Instr {
  uint8_t op;
  uint8_t flags;
  int16_t in[4];    // wire slot per input, -1 = not connected (None)
  int16_t out[3];   // wire slot per output port, -1 = nobody reads it (not materialised)
  uint16_t state;   // first per-voice state word
  uint16_t param;   // first per-voice parameter word
  uint16_t aux;     // ring id / module index (noise key) / first channel (output) / table offset (sequencers)
  uint8_t warp;     // warp of the group that executes this instruction
  uint8_t stage;    // pipeline delay in chunks
  float imm;        // oscillator / ADSR sample rate
  uint8_t n_ch;     // OUTPUT / MIX: channels covered by this instruction (1..4); sequencers: steps (1..64);
                    // OSC: 1 = delta arrives on wires in[2] (lo) / in[3] (hi) from an OP_OSC_DELTA
  uint8_t pad[3];
};
static_assert(sizeof(Instr) == 32, "Instr must stay 32 bytes (staged to shared memory as uint4 pairs)");

// One wire slot: a ring of (mask + 1) tiles of [K samples][32 voices] f32, first tile `base`.
struct WireDesc {
  uint16_t base;
  uint16_t mask;
};

// A Sample module's table as the device sees it: four words in Program::tables (uniform over voices).
// The samples themselves stay in HBM (RenderArgs::waves), all modules' tables back to back.
struct WaveDesc {
  uint32_t offset;  // first sample inside the concatenated wave buffer
  uint32_t len;     // samples (< 2^31)
  float ratio;      // wavebox.sample_rate / self.sample_rate, the f32 quotient (sample.rs:234)
  uint32_t is_new;  // WaveBox.new (sample.rs:212-216): rewind at the start of this render
};
static_assert(sizeof(WaveDesc) == 16, "WaveDesc is four table words");

// Per-voice state words (u32 slots, SoA [word][voice] in HBM)
//   OSC   : pos (f64, 2 words), sync_last                         oscillator.rs:21,23
//   NOISE : sample counter (u64, 2 words)
//   MOOG  : f, p, q, b[0..4], freq, res                            filter.rs:48-56
//   ADSR  : phase, r_val, from_a_val, mode | gate_last << 8        adsr.rs:14-21
//   GRIDSEQ : current_step | step_last << 16 | sync_last << 17, last cv   sequencer.rs:24-27
//   PATSEQ  : current_step | step_last << 16 | sync_last << 17 (one copy per instruction)
//   SAMPLE  : pos (f32), playing | gate_last << 1                    sample.rs:78-82
constexpr int kStateOsc = 3, kStateNoise = 2, kStateMoog = 10, kStateAdsr = 4, kStateGridSeq = 2, kStatePatSeq = 1, kStateSample = 2;
// Per-voice parameter words (SoA [word][voice] in HBM)
//   OSC   : val, delta (f64, 2 words; host-computed 440*2^val/sr, used when CV is None)
//   MOOG  : freq, res, exp_amt        ADSR : a_sec, d_sec, s_val, r_sec
//   MIXER : gain[0..3]                MATH : constant
constexpr int kParamOsc = 3, kParamMoog = 3, kParamAdsr = 4, kParamMixer = 4, kParamMath = 1;

// Where a parameter word comes from when the table is (re)built for a voice range.
struct ParamSource {
  int module;       // index into the patch's module list at plan time
  int pid;          // srk_param id; -1 => derived oscillator delta (lo word), -2 => (hi word)
};

struct Program {
  std::vector<Instr> code;             // sorted by (warp, plan order), terminated by OP_END
  std::vector<uint16_t> warp_begin;    // n_warps + 1 offsets into `code`
  std::vector<WireDesc> wires;         // one per wire slot
  std::vector<int32_t> tables;         // sequencer step tables / WaveDescs (uniform over voices), Instr::aux indexes it
  std::vector<int> wave_modules;       // Sample modules (patch list index) in the order their tables are concatenated
  uint64_t wave_total = 0;             // samples in the concatenated wave buffer
  std::vector<uint32_t> state_init;    // one initial value per state word
  std::vector<ParamSource> param_src;  // one per parameter word
  uint32_t n_tiles = 0;                // wire tiles per group
  uint32_t n_warps = 1;                // S
  uint32_t n_stages = 1;               // max stage + 1
  uint32_t max_ring_store_stage = 0;
  uint32_t max_cost = 0, sum_cost = 0; // cost model (program.cpp): slowest scheduled instruction / all modules in series
  uint32_t n_rings = 0;
  uint32_t channels = 0;
  uint32_t ring_len = 0;               // buffer_size
};

}  // namespace srk
