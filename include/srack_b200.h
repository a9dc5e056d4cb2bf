/* ============================================================================
 * srack_b200.h -- C ABI of the B200-native s-rack voice renderer.
 *
 * Drop-in boundary for the reference's module-graph tick (sharph/s-rack,
 * src/synth.rs + src/synth/{oscillator,filter,adsr,vca,mixer,math,sequencer,sample,output}.rs).
 * The reference has no FFI of its own: its seam is the Rust `SynthModule`
 * trait plus the free functions `plan_execution`, `execute`, `get_catalog`.
 * Every entry point below names the reference interface (file:line) it
 * replaces; INTEGRATION.md shows the Rust binding a maintainer would add.
 *
 * Conventions
 *   - plain C: opaque handles, pointers and sizes only; no C++/torch types.
 *   - every function returns an `int` status (SRK_OK == 0) unless documented
 *     otherwise; the reference's `Err(())` maps to SRK_ERR_PORT, its
 *     `.unwrap()` panics map to error codes -- nothing aborts across the ABI.
 *   - a patch handle is single-threaded (the caller serialises, as the
 *     reference does with `Mutex<plan>`, src/main.rs:60); distinct patches may
 *     be used from distinct threads.
 *   - the library owns modules, per-voice state and all device memory; the
 *     caller owns output buffers (host by default, device with
 *     SRK_RENDER_DEVICE_OUT).
 *   - there is NO CPU fallback: rendering without a CUDA device fails with
 *     SRK_ERR_NO_DEVICE.
 * ==========================================================================*/
#ifndef SRACK_B200_H
#define SRACK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SRK_API __attribute__((visibility("default")))
#else
#define SRK_API
#endif

/* ---- status codes ------------------------------------------------------- */
enum srk_status {
  SRK_OK = 0,
  SRK_ERR_ARG = 1,         /* NULL / foreign handle, bad argument                 */
  SRK_ERR_PORT = 2,        /* port index out of range: the reference's Err(())    */
  SRK_ERR_KIND = 3,        /* unknown module kind / catalog name                  */
  SRK_ERR_UNSUPPORTED = 4, /* catalog entry outside the hot path  */
  SRK_ERR_PARAM = 5,       /* unknown parameter id for this kind                  */
  SRK_ERR_SELF_LOOP = 6,   /* module wired to itself (deadlocks in the reference) */
  SRK_ERR_NO_OUTPUT = 7,   /* patch has no Output module (ui.rs:84-96 -> empty plan) */
  SRK_ERR_NOT_PLANNED = 8, /* wiring changed since the last srk_plan()            */
  SRK_ERR_SIZE = 9,        /* per-voice array shorter than the voices rendered... */
  SRK_ERR_NO_DEVICE = 10,  /* no CUDA device: the product has no CPU path         */
  SRK_ERR_CUDA = 11,       /* CUDA runtime failure, see srk_last_error()          */
  SRK_ERR_LIMIT = 12       /* patch too large for one thread block's shared memory */
};

/* ---- module kinds: the reference catalog (src/synth.rs:421-515) + Output,
 *      which the app creates itself (src/main.rs:130) -------------------- */
enum srk_kind {
  SRK_KIND_OUTPUT = 0,      /* src/synth/output.rs      "Output"      */
  SRK_KIND_OSCILLATOR = 1,  /* src/synth/oscillator.rs  "Oscillator"  */
  SRK_KIND_NOISE = 2,       /* src/synth/oscillator.rs  "Noise"       */
  SRK_KIND_ADSR = 3,        /* src/synth/adsr.rs        "ADSR"        */
  SRK_KIND_VCA = 4,         /* src/synth/vca.rs         "VCA"         */
  SRK_KIND_MOOG_FILTER = 5, /* src/synth/filter.rs      "Moog Filter" */
  SRK_KIND_MONO_MIXER = 6,  /* src/synth/mixer.rs       "Mono Mixer"  */
  SRK_KIND_ADD = 7,         /* src/synth/math.rs        "Add"         */
  SRK_KIND_SUBTRACT = 8,    /* src/synth/math.rs        "Subtract"    */
  SRK_KIND_MULTIPLY = 9,    /* src/synth/math.rs        "Multiply"    */
  SRK_KIND_NON_LINEAR = 10, /* src/synth/math.rs        "Non-Linear"  */
  SRK_KIND_GRID_SEQUENCER = 11,    /* src/synth/sequencer.rs  "Grid Sequencer"    */
  SRK_KIND_PATTERN_SEQUENCER = 12, /* src/synth/sequencer.rs  "Pattern Sequencer" */
  SRK_KIND_SAMPLE = 13,            /* src/synth/sample.rs     "Sample"            */
  SRK_KIND_COUNT = 14
};

/* ---- parameter ids (the reference mutates struct fields from ui(); there
 *      is no generic setter -- these enumerate those fields) -------------- */
enum srk_param {
  /* Oscillator: oscillator.rs:12 (val, UI range -9..6 at :221), :22 antialiasing */
  SRK_OSC_VAL = 0,
  SRK_OSC_ANTIALIASING = 1, /* 0 / 1, uniform only */
  /* ADSR: adsr.rs:10-13, defaults :39-42 */
  SRK_ADSR_A_SEC = 0,
  SRK_ADSR_D_SEC = 1,
  SRK_ADSR_S_VAL = 2,
  SRK_ADSR_R_SEC = 3,
  /* VCA: vca.rs:14 */
  SRK_VCA_NEGATIVE = 0, /* 0 / 1, uniform only */
  /* Moog filter: filter.rs:21-23, defaults :36-38 */
  SRK_MOOG_FREQ = 0,
  SRK_MOOG_RES = 1,
  SRK_MOOG_EXP_AMT = 2,
  /* Mono mixer: mixer.rs:11, gains of inputs 0..3 */
  SRK_MIXER_GAIN0 = 0,
  SRK_MIXER_GAIN1 = 1,
  SRK_MIXER_GAIN2 = 2,
  SRK_MIXER_GAIN3 = 3,
  /* Add/Subtract/Multiply/Non-Linear: math.rs:21,184 */
  SRK_MATH_CONSTANT = 0,
  /* Grid sequencer: sequencer.rs:20 steps_per_octave (u16, default 12 at :44), uniform only */
  SRK_GRIDSEQ_STEPS_PER_OCTAVE = 0
};

/* Sequence cells (srk_set_sequence).  Grid sequencer, sequencer.rs:18 `Vec<Option<(u16, bool)>>`:
 * SRK_SEQ_NONE, or SRK_GRID_CELL(val, hold).  Pattern sequencer, sequencer.rs:341
 * `Vec<Vec<Option<bool>>>`: SRK_SEQ_NONE, 0 (Some(false): pass the clock) or 1 (Some(true): hold). */
#define SRK_SEQ_NONE (-1)
#define SRK_GRID_CELL(val, hold) ((int32_t)(((val) & 0xFFFF) | ((hold) ? 0x10000 : 0)))
#define SRK_SEQ_MAX_STEPS 64
#define SRK_PATTERN_ROWS 8

/* ---- render flags ------------------------------------------------------- */
enum srk_render_flags {
  SRK_RENDER_DEVICE_OUT = 1u << 0, /* `stems` / `mix` are device pointers on the patch's device */
  SRK_RENDER_ASYNC = 1u << 1       /* return after enqueueing (device outputs only); srk_sync() waits */
};

/* src/synth.rs:20-25 `AudioConfig { sample_rate: u16, buffer_size: usize, channels: u8 }` */
typedef struct srk_audio_config {
  uint16_t sample_rate;
  size_t buffer_size;
  uint8_t channels;
} srk_audio_config;

typedef struct srk_patch srk_patch;   /* the module list + plan (ui.rs:51-60 SynthModuleWorkspaceImpl) */
typedef struct srk_module srk_module; /* one SharedSynthModule (src/synth.rs:270); owned by its patch */

/* ---- library ------------------------------------------------------------ */
SRK_API const char* srk_version(void);
SRK_API const char* srk_status_string(int status);

/* ---- catalog: get_catalog(), src/synth.rs:421-515 ----------------------- */
/* Number of catalog entries (the reference's 14, in its order, then "Output"). */
SRK_API int srk_catalog_size(void);
/* Name of entry i, e.g. "Oscillator"; NULL when i is out of range. */
SRK_API const char* srk_catalog_name(int i);
/* srk_kind of entry i, or -1 for entries outside the hot path ("Freeverb"). */
SRK_API int srk_catalog_kind(int i);

/* ---- patch lifetime ----------------------------------------------------- */
/* Replaces the workspace that owns `modules` and `plan` (ui.rs:51-60, main.rs:103-125). */
SRK_API int srk_patch_create(const srk_audio_config* cfg, srk_patch** out);
SRK_API void srk_patch_destroy(srk_patch* patch);
/* SynthModule::set_audio_config for every module (synth.rs:261; ui.rs set_audio_config).
 * As in the reference, ADSR keeps the sample rate it was built with (adsr.rs:69-71)
 * and Output drops its connections (output.rs:40-45).  Resets all voice state. */
SRK_API int srk_set_audio_config(srk_patch* patch, const srk_audio_config* cfg);
SRK_API int srk_get_audio_config(const srk_patch* patch, srk_audio_config* out);
/* Seed of the counter-based noise generator that stands in for the reference's
 * unseeded rand::random (oscillator.rs:385).  Default 0x5EED5EED. */
SRK_API int srk_set_seed(srk_patch* patch, uint64_t seed);
/* CUDA device ordinal used by this patch (default: current device at first render). */
SRK_API int srk_set_device(srk_patch* patch, int device);
SRK_API const char* srk_last_error(const srk_patch* patch);

/* ---- modules: X::new(&AudioConfig) via the catalog closures (synth.rs:421-515,
 *      main.rs:130,151-159); appended to the patch's module list ----------- */
SRK_API int srk_module_create(srk_patch* patch, int kind, srk_module** out);
SRK_API int srk_module_create_by_name(srk_patch* patch, const char* catalog_name, srk_module** out);
/* ui.rs delete_module: disconnects every input fed by `module`, removes it from the list. */
SRK_API int srk_module_remove(srk_patch* patch, srk_module* module);
SRK_API size_t srk_module_count(const srk_patch* patch);
SRK_API srk_module* srk_module_at(const srk_patch* patch, size_t index);

/* SynthModule getters, src/synth.rs:223-232 */
SRK_API const char* srk_get_id(const srk_module* m);   /* uuid-v4 text, valid until the module is removed */
SRK_API const char* srk_get_name(const srk_module* m); /* "Oscillator", ... */
SRK_API int srk_get_kind(const srk_module* m);         /* srk_kind, -1 on NULL */
SRK_API int srk_get_num_inputs(const srk_module* m);   /* u8 in the reference; -1 on NULL */
SRK_API int srk_get_num_outputs(const srk_module* m);
/* *label is NULL for the reference's Ok(None); SRK_ERR_PORT for its Err(()). */
SRK_API int srk_get_input_label(const srk_module* m, uint8_t input_idx, const char** label);
SRK_API int srk_get_output_label(const srk_module* m, uint8_t output_idx, const char** label);

/* set_input / disconnect_input / disconnect_inputs / get_input, src/synth.rs:228,234-246.
 * Unlike the reference the source port is validated here (SRK_ERR_PORT) instead of
 * panicking later in resolve_input (synth.rs:251). */
SRK_API int srk_connect(srk_module* sink, uint8_t input_idx, srk_module* src, uint8_t src_port);
SRK_API int srk_disconnect(srk_module* sink, uint8_t input_idx);
SRK_API int srk_disconnect_inputs(srk_module* sink);
/* *src is NULL when the input is not connected. */
SRK_API int srk_get_input(const srk_module* sink, uint8_t input_idx, srk_module** src, uint8_t* src_port);

/* parameters (struct fields set from ui(): oscillator.rs:12,221; filter.rs:226-236;
 * adsr.rs:223-257; mixer.rs:126-131; math.rs:165,318) */
SRK_API int srk_set_param_f32(srk_module* m, int param_id, float value);
SRK_API int srk_get_param_f32(const srk_module* m, int param_id, float* value);
/* New axis: one value per voice (global voice index), e.g. detune.  The array is copied. */
SRK_API int srk_set_param_f32_per_voice(srk_module* m, int param_id, const float* values, size_t n_voices);

/* The step table a sequencer's ui() edits (sequencer.rs:98-188 grid, :388-470 pattern), the same for
 * every voice.  Grid sequencer: `cells[n_steps]`; pattern sequencer: `cells[8][n_steps]` row major;
 * 1 <= n_steps <= 64 (the reference's UI limits).  Default: 64 steps of None.  Voice state is kept
 * (the reference edits the table under the module lock while the audio thread keeps running). */
SRK_API int srk_set_sequence(srk_module* m, const int32_t* cells, size_t n_steps);
/* *n_steps receives the current length; up to `cap` cells are copied out (rows x steps for the pattern). */
SRK_API int srk_get_sequence(const srk_module* m, int32_t* cells, size_t cap, size_t* n_steps);

/* ---- Sample module: the WaveBox (sample.rs:14-20) behind "Load Sample..." (sample.rs:242-257).
 * One table per module, shared by every voice (each voice has its own play position).  As in the
 * reference (`new = true`, sample.rs:66,212-216) a load rewinds every voice at the start of the
 * next render: position 0, not playing; the gate detector keeps its state. ------------------- */
/* WaveBox::load, sample.rs:32-69 (WAV parsing as hound 3.5.1 does it): RIFF/WAVE, PCM 8/16/24 bit
 * or 32-bit float, channel 0 only, converted as sample.rs:49-54.  Malformed header: SRK_ERR_ARG,
 * the module keeps what it had.  32-bit integer PCM or a truncated data chunk: SRK_ERR_UNSUPPORTED
 * and -- like the reference, which has already run `samples.clear()` by then -- an EMPTY table. */
SRK_API int srk_load_wav(srk_module* m, const void* wav_bytes, size_t n_bytes);
/* The decoded state directly: `samples[n]` (copied) and the file's sample rate. */
SRK_API int srk_set_sample(srk_module* m, const float* samples, size_t n, float sample_rate);
/* *n receives the table length; up to `cap` samples are copied out. */
SRK_API int srk_get_sample(const srk_module* m, float* samples, size_t cap, size_t* n, float* sample_rate);

/* ---- WAV export of a render (the reference only reads WAV).  `planar` is
 * [channels][n_samples] f32 as srk_render() writes `mix`; bits = 32 -> IEEE float, 16 / 24 ->
 * PCM (x * 2^(bits-1), rounded to nearest, clamped).  Host-only, no device involved. --------- */
SRK_API int srk_write_wav(const char* path, const float* planar, unsigned channels, size_t n_samples,
                          uint32_t sample_rate, int bits);

/* ---- .srk patch files: FileFormat (src/ui.rs:578-586) in MessagePack as rmp-serde 1.3.0 writes it
 *      (ui.rs:98-114 serialize, :115-134 deserialize).
 * Load empties the patch and rebuilds it: module list in the order the reference ends up with (the
 * file's reversed, ui.rs:652-660), ids, parameters, sequencer tables, Sample tables, connections
 * (entries naming unknown ids or bad ports are skipped like the reference's `let _ = set_input(..)`,
 * ui.rs:662-681; their count goes to *n_skipped_connections when non-NULL) and the DSP state the
 * modules were saved with (phase, filter memory, envelope stage, step counters, play position,
 * detectors): every voice starts from it and srk_reset() returns to it, like the reference's
 * deserialized modules carry on from where they were saved.  Serialized port buffers are not imported
 * (a wire the cycle breaker cut starts with an empty history).  A file that contains a
 * Freeverb module is refused with SRK_ERR_UNSUPPORTED and the patch is left as it was; a malformed file
 * gives SRK_ERR_ARG.  Per-voice parameter arrays are dropped (the file has one value per field).
 * Call srk_plan() afterwards (the reference plans at the end of deserialize). ------------------- */
SRK_API int srk_patch_load_srk(srk_patch* patch, const void* bytes, size_t n_bytes, size_t* n_skipped_connections);
/* The patch as the reference would save it: zeroed port buffers of buffer_size samples, X::new() DSP state
 * for modules built through this ABI, the loaded state for modules that came from a file (the device state
 * of a render is per voice and is not written back).  *bytes stays valid until the next save or patch destroy. */
SRK_API int srk_patch_save_srk(srk_patch* patch, const void** bytes, size_t* n_bytes);

/* ---- planning: plan_execution(output, &all_modules, &mut plan), src/synth.rs:128-212,
 *      called as in ui.rs:63-82 (output = first Output in the module list) -------- */
SRK_API int srk_plan(srk_patch* patch);
/* Plan order; `cap` entries available in `out`, *n receives the plan length. */
SRK_API int srk_plan_get(const srk_patch* patch, srk_module** out, size_t cap, size_t* n);
/* Wires removed by the cycle breaker (synth.rs:168-192) as (reader, writer) pairs. */
SRK_API int srk_plan_cuts(const srk_patch* patch, srk_module** readers, srk_module** writers, size_t cap, size_t* n);
/* Reorders the module list (`all_modules`), as the reference test's shuffle does
 * (synth.rs:561-567).  `order` must be a permutation of the patch's modules. */
SRK_API int srk_set_module_order(srk_patch* patch, srk_module* const* order, size_t n);

/* ---- rendering: execute(&plan) (src/synth.rs:97-101, called from main.rs:62-63)
 *      + reading OutputModule.bufs (main.rs:64-75), for n_voices instances of the
 *      patch at once.  Voices [voice_offset, voice_offset + n_voices) of the global
 *      voice axis are rendered for n_samples samples; all module state persists
 *      across calls exactly as the reference's modules persist across blocks.
 *        stems: [channels][n_samples][n_voices] f32, or NULL
 *        mix:   [channels][n_samples] f32 (sum over the rendered voices), or NULL
 *      Host pointers unless SRK_RENDER_DEVICE_OUT.  Changing n_voices/voice_offset
 *      between calls resets the voice state. ------------------------------- */
SRK_API int srk_render(srk_patch* patch, size_t n_voices, size_t voice_offset, size_t n_samples,
                       unsigned flags, float* stems, float* mix);
/* Same, enqueued on a caller-supplied cudaStream_t (passed as void*). */
SRK_API int srk_render_on_stream(srk_patch* patch, size_t n_voices, size_t voice_offset, size_t n_samples,
                                 unsigned flags, float* stems, float* mix, void* cuda_stream);
/* How often the voice state of this patch has been (re)initialised so far: by the first render, srk_reset(), a
 * re-plan after a wiring change, srk_state_import() -- and IMPLICITLY whenever srk_render() is called with another
 * n_voices or voice_offset than the call before (the state belongs to one voice range; alternate ranges through
 * separate patches, or carry them with srk_state_export / srk_state_import).  A caller that streams a render in blocks
 * can assert that the number does not move. */
SRK_API uint64_t srk_state_epoch(const srk_patch* patch);
SRK_API int srk_sync(srk_patch* patch);
/* Back to X::new() state -- or the state a loaded .srk file carried -- for every module of every voice
 * (and empty feedback history). */
SRK_API int srk_reset(srk_patch* patch);

/* ---- per-voice DSP state out and in: the device-side analogue of the state every reference module serializes
 *      (`#[derive(Serialize)]` on OscillatorModule.pos, InternalMoogFilterState, ADSRModule.phase/mode/r_val ...,
 *      written by ui.rs:98-114 and restored by ui.rs:115-134) -- a checkpoint of a render in progress.
 * Export: every voice's state words, the history of the wires the cycle breaker cut and the absolute sample index of
 * the range last rendered on this patch, as one opaque, versioned blob (valid until the next export or patch destroy;
 * SRK_ERR_ARG when nothing has been rendered since the last wiring change).
 * Import: into a planned patch with the same graph (module kinds in plan order, cut wires, buffer_size -- else
 * SRK_ERR_ARG; SRK_ERR_SIZE for a truncated blob); the next srk_render() of the blob's n_voices / voice_offset
 * continues bit for bit where the exporting patch stopped.  Parameters, sequencer tables and Sample tables are not
 * part of the blob (they belong to the patch description / the .srk file). */
SRK_API int srk_state_export(srk_patch* patch, const void** blob, size_t* n_bytes);
SRK_API int srk_state_import(srk_patch* patch, const void* blob, size_t n_bytes);

/* Scheduling hint: `n_voices` voices of OTHER patches are rendered on this patch's device at the same time (several
 * patches, each on its own stream).  Launch shapes are chosen from the voice groups an SM has to hold, so concurrent
 * renders should be scheduled for the sum.  Default 0.  Changing it re-plans the launch and resets the voice state. */
SRK_API int srk_set_co_resident_voices(srk_patch* patch, size_t n_voices);

/* ---- instrumentation ---------------------------------------------------- */
/* Device time (CUDA events on the render stream) of the voice kernel alone and of the
 * whole call, for the last completed render; kernel launches issued so far. */
SRK_API int srk_last_render_ms(srk_patch* patch, float* kernel_ms, float* total_ms);
SRK_API uint64_t srk_launch_count(const srk_patch* patch);
/* Compiled-program facts for the last plan and `n_voices`: samples per chunk, threads per
 * block, shared-memory bytes per block, wire slots, state words and parameter words per
 * voice, feedback rings (delayed wires), warps per 32-voice group, pipeline stages,
 * wire tiles per group and voice groups per thread block (> 1 only in the one-warp schedule,
 * where block_threads = 32 * groups_per_block and smem_bytes covers all of them). */
typedef struct srk_program_info {
  uint32_t n_instr, step_samples, block_threads, smem_bytes;
  uint32_t n_wires, state_words, param_words, n_rings;
  uint32_t n_warps, n_stages, n_tiles, groups_per_block;
  /* fused = 1: the launch uses the patch-specialised kernel (wires in registers, step_samples = the 32-sample
   * output tile, groups_per_block = independent warps per block); registers per thread and local (spill) bytes
   * are known once the kernel has been loaded by a render on this patch, else 0. */
  uint32_t fused, fused_group, fused_regs, fused_local_bytes;
} srk_program_info;
SRK_API int srk_get_program_info(srk_patch* patch, size_t n_voices, srk_program_info* out);
/* The CUDA C++ translation unit generated for this patch (the wiring; the DSP is csrc/fused_ops.cuh) when a
 * render of n_voices would use a fused kernel, else an empty string.  Valid until the next call on the patch. */
SRK_API int srk_fused_source(srk_patch* patch, size_t n_voices, const char** source, size_t* n_bytes);
/* Compiles that kernel -- and the alternative launch shapes a long render measures against it (srk_schedule_report) --
 * for sm_100a into the on-disk cubin cache (kernel_cache/ next to the library, or $SRK_KERNEL_CACHE) so that the first
 * render does not pay for NVRTC.  Needs no GPU.  *compiled = the number of kernels compiled now (0: all cached already,
 * or the launch would not use a fused kernel).  NVRTC is the CUDA toolkit's libnvrtc.so.12, loaded by path
 * (/usr/local/cuda/lib64; $SRK_NVRTC_LIB overrides); when $SRK_NVRTC_ALT names a copy of ANOTHER version every
 * candidate is built by both and the measurement picks between them as well ($SRK_NVRTC_ALT=0: never). */
SRK_API int srk_precompile(srk_patch* patch, size_t n_voices, int* compiled);
/* Identity of the kernel image a render of n_voices would launch: "fused:<hash of generated source + op headers +
 * compiler options + NVRTC version>" or "interpreter:<hash of the kernel sources at build time>:<pipelined|solo|solo_full>".  Profiles
 * are stamped with it.  Valid until the next call on the patch. */
SRK_API int srk_kernel_id(srk_patch* patch, size_t n_voices, const char** id);
/* How the launch shape in use was chosen.  The first render of at least 16384 samples after a (re)plan measures the
 * plausible launch shapes (fused kernel with 4 or 8 samples per straight-line group, one more pipeline stage, the
 * interpreter's pipeline; each fused one as either NVRTC version builds it) for a few thousand samples each on a scratch copy of the voice state and keeps the fastest;
 * every shape computes the same bits.  *report: "" before that, else the winner and the measured times (or the cached
 * decision, kernel_cache/<key>.tune).  SRK_TUNE=0 in the environment, or any forced schedule knob, disables it.
 * srk_get_program_info / srk_kernel_id describe the shape in use once a render has happened.  Valid until the next call. */
SRK_API int srk_schedule_report(srk_patch* patch, const char** report);
/* The compiled, scheduled device program itself (what execute() becomes for n_voices voices):
 * one entry per instruction -- the modules of the plan in plan order (src/synth.rs:97-101), plus
 * ring loads/stores for the wires the cycle breaker cut (synth.rs:168-192), the stems / mixdown
 * instructions of the Output, time-split oscillator copies -- sorted by the warp that runs it,
 * terminated by an op-0 entry; and one entry per wire slot.  For inspection and tests: the
 * schedule is checked on the CPU without a device (tests/test_program.py). */
typedef struct srk_instr_info {
  uint8_t op;      /* 0 end, 1 ring load, 2 ring store, 3 oscillator, 4 noise, 5 moog, 6 adsr, 7 vca,
                      8 mixer, 9 math, 10 output (stems), 11 mix, 12 moog coefficients,
                      13 grid sequencer, 14 pattern sequencer (three output ports per instruction),
                      15 oscillator V/oct conversion (delta = 440 * 2^cv / sr on a wire pair),
                      16 sample player, 17 / 18 oscillator phase recurrence / stateless shaping (experimental split) */
  uint8_t flags;   /* math: operation; oscillator: (copies << 4) | copy index when time-split, bit 3 = antialiasing off;
                      VCA: bit 0 = negative */
  uint8_t warp;    /* warp of the 32-voice group that executes it */
  uint8_t stage;   /* works on chunk (iteration - stage) */
  int16_t in[4];   /* wire slot per input, -1 = not connected */
  int16_t out[3];  /* wire slot per output port, -1 = nobody reads it */
  uint8_t n_ch;    /* output / mix: channels covered */
  uint8_t reserved;
  uint16_t state, param, aux; /* first state word, first parameter word, ring / channel / module index */
} srk_instr_info;
typedef struct srk_wire_info {
  uint16_t first_tile, n_tiles; /* ring of n_tiles [chunk][32 voices] tiles; chunk c lives in tile c % n_tiles */
} srk_wire_info;
SRK_API int srk_get_program(srk_patch* patch, size_t n_voices, srk_instr_info* instrs, size_t instr_cap,
                            size_t* n_instr, srk_wire_info* wires, size_t wire_cap, size_t* n_wires);

#ifdef __cplusplus
}
#endif
#endif /* SRACK_B200_H */
