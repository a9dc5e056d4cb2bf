// extern "C" surface declared in include/srack_b200.h.  Pure host code: argument
// checking, the module list and wiring; everything that touches the GPU is in
// engine.cu.  Nothing here throws across the ABI.
#include <algorithm>
#include <cstring>
#include <exception>
#include <new>

#include "engine.hpp"
#include "patch.hpp"
#include "srkfile.hpp"
#include "wav.hpp"

namespace {

// get_catalog() order, src/synth.rs:421-515, then "Output" (created by the app, main.rs:130).
struct CatalogEntry { const char* name; int kind; };
const CatalogEntry kCatalog[] = {
    {"Oscillator", SRK_KIND_OSCILLATOR}, {"Noise", SRK_KIND_NOISE},
    {"Grid Sequencer", SRK_KIND_GRID_SEQUENCER}, {"Pattern Sequencer", SRK_KIND_PATTERN_SEQUENCER},
    {"ADSR", SRK_KIND_ADSR},             {"VCA", SRK_KIND_VCA},
    {"Moog Filter", SRK_KIND_MOOG_FILTER}, {"Mono Mixer", SRK_KIND_MONO_MIXER},
    {"Sample", SRK_KIND_SAMPLE},        {"Add", SRK_KIND_ADD},
    {"Subtract", SRK_KIND_SUBTRACT},     {"Multiply", SRK_KIND_MULTIPLY},
    {"Non-Linear", SRK_KIND_NON_LINEAR}, {"Freeverb", -1},
    {"Output", SRK_KIND_OUTPUT},
};
constexpr int kCatalogSize = (int)(sizeof(kCatalog) / sizeof(kCatalog[0]));

int fail(srk_patch* p, int code, const char* msg) {
  if (p) p->last_error = msg;
  return code;
}

void touch_wiring(srk_patch* p) {
  ++p->wiring_epoch;
  p->planned = false;
}

// Nothing may throw across the ABI: allocation failures (a huge per-voice array, a hostile .srk file) become a status.
template <class F>
int guarded(srk_patch* p, F&& f) noexcept {
  try {
    return f();
  } catch (const std::bad_alloc&) {
    return fail(p, SRK_ERR_LIMIT, "out of host memory");
  } catch (const std::exception& e) {
    if (p) {
      try { p->last_error = e.what(); } catch (...) {}
    }
    return SRK_ERR_LIMIT;
  } catch (...) {
    return fail(p, SRK_ERR_LIMIT, "unexpected exception");
  }
}

int n_inputs_for(const srk_patch* p, int kind) {
  int n = srk::kind_info(kind).n_inputs;
  return n < 0 ? p->cfg.channels : n;
}

}  // namespace

extern "C" {

const char* srk_version(void) { return "srack_b200 0.1 (sm_100a; s-rack module-graph tick @20e549b)"; }

const char* srk_status_string(int s) {
  switch (s) {
    case SRK_OK: return "ok";
    case SRK_ERR_ARG: return "bad argument";
    case SRK_ERR_PORT: return "port index out of range";
    case SRK_ERR_KIND: return "unknown module kind";
    case SRK_ERR_UNSUPPORTED: return "module kind outside the hot path";
    case SRK_ERR_PARAM: return "unknown parameter id";
    case SRK_ERR_SELF_LOOP: return "module wired to itself";
    case SRK_ERR_NO_OUTPUT: return "patch has no Output module";
    case SRK_ERR_NOT_PLANNED: return "patch not planned";
    case SRK_ERR_SIZE: return "per-voice array too short";
    case SRK_ERR_NO_DEVICE: return "no CUDA device (no CPU path)";
    case SRK_ERR_CUDA: return "CUDA error";
    case SRK_ERR_LIMIT: return "patch exceeds device limits";
  }
  return "unknown status";
}

int srk_catalog_size(void) { return kCatalogSize; }
const char* srk_catalog_name(int i) { return i >= 0 && i < kCatalogSize ? kCatalog[i].name : nullptr; }
int srk_catalog_kind(int i) { return i >= 0 && i < kCatalogSize ? kCatalog[i].kind : -1; }

int srk_patch_create(const srk_audio_config* cfg, srk_patch** out) {
  if (!cfg || !out) return SRK_ERR_ARG;
  if (cfg->buffer_size == 0 || cfg->sample_rate == 0) return SRK_ERR_ARG;
  srk_patch* p = new (std::nothrow) srk_patch();
  if (!p) return SRK_ERR_ARG;
  p->cfg = *cfg;
  *out = p;
  return SRK_OK;
}

void srk_patch_destroy(srk_patch* patch) { delete patch; }

int srk_set_audio_config(srk_patch* p, const srk_audio_config* cfg) {
  if (!p || !cfg) return SRK_ERR_ARG;
  if (cfg->buffer_size == 0 || cfg->sample_rate == 0) return fail(p, SRK_ERR_ARG, "zero buffer_size / sample_rate");
  p->cfg = *cfg;
  for (srk_module* m : p->modules) {
    if (m->kind == SRK_KIND_OSCILLATOR || m->kind == SRK_KIND_SAMPLE) m->osc_sample_rate = cfg->sample_rate;  // oscillator.rs:83-84, sample.rs:120-123
    // ADSR keeps its construction-time sample rate (adsr.rs:69-71)
    if (m->kind == SRK_KIND_OUTPUT) m->inputs.assign(cfg->channels, {nullptr, 0});  // output.rs:40-45
  }
  touch_wiring(p);
  srk::engine_invalidate_state(p);
  return SRK_OK;
}

int srk_get_audio_config(const srk_patch* p, srk_audio_config* out) {
  if (!p || !out) return SRK_ERR_ARG;
  *out = p->cfg;
  return SRK_OK;
}

int srk_set_seed(srk_patch* p, uint64_t seed) {
  if (!p) return SRK_ERR_ARG;
  p->seed = seed;
  return SRK_OK;
}

int srk_set_device(srk_patch* p, int device) {
  if (!p || device < 0) return SRK_ERR_ARG;
  if (p->engine) return fail(p, SRK_ERR_ARG, "device already in use by this patch");
  p->device = device;
  return SRK_OK;
}

const char* srk_last_error(const srk_patch* p) { return p ? p->last_error.c_str() : "null patch"; }

int srk_module_create(srk_patch* p, int kind, srk_module** out) { return guarded(p, [&]() -> int {
  if (!p || !out) return SRK_ERR_ARG;
  if (kind < 0 || kind >= SRK_KIND_COUNT) return fail(p, SRK_ERR_KIND, "unknown module kind");
  auto m = std::make_unique<srk_module>();
  m->patch = p;
  m->kind = kind;
  m->id = srk::make_uuid_v4();
  m->inputs.assign(n_inputs_for(p, kind), {nullptr, 0});
  const srk::KindInfo& ki = srk::kind_info(kind);
  for (int i = 0; i < ki.n_params; ++i) m->param[i] = ki.param_default[i];
  m->osc_sample_rate = p->cfg.sample_rate;
  m->adsr_sample_rate = (float)p->cfg.sample_rate;
  if (kind == SRK_KIND_GRID_SEQUENCER || kind == SRK_KIND_PATTERN_SEQUENCER) {  // vec![None; 64] (sequencer.rs:40,360)
    m->seq_steps = SRK_SEQ_MAX_STEPS;
    m->sequence.assign((kind == SRK_KIND_PATTERN_SEQUENCER ? SRK_PATTERN_ROWS : 1) * m->seq_steps, SRK_SEQ_NONE);
  }
  *out = m.get();
  p->modules.push_back(m.get());
  p->owned.push_back(std::move(m));
  touch_wiring(p);
  return SRK_OK;
}); }

int srk_module_create_by_name(srk_patch* p, const char* name, srk_module** out) { return guarded(p, [&]() -> int {
  if (!p || !name || !out) return SRK_ERR_ARG;
  for (const auto& e : kCatalog)
    if (std::strcmp(e.name, name) == 0) {
      if (e.kind < 0) return fail(p, SRK_ERR_UNSUPPORTED, "catalog entry is outside the hot path");
      return srk_module_create(p, e.kind, out);
    }
  return fail(p, SRK_ERR_KIND, "no such catalog entry");
}); }

int srk_module_remove(srk_patch* p, srk_module* m) {
  if (!p || !m || m->patch != p) return SRK_ERR_ARG;
  for (srk_module* other : p->modules)
    for (auto& in : other->inputs)
      if (in.first == m) in = {nullptr, 0};
  p->modules.erase(std::remove(p->modules.begin(), p->modules.end(), m), p->modules.end());
  p->owned.erase(std::remove_if(p->owned.begin(), p->owned.end(), [&](auto& u) { return u.get() == m; }), p->owned.end());
  p->plan.clear();
  p->cuts.clear();
  touch_wiring(p);
  return SRK_OK;
}

size_t srk_module_count(const srk_patch* p) { return p ? p->modules.size() : 0; }
srk_module* srk_module_at(const srk_patch* p, size_t i) { return p && i < p->modules.size() ? p->modules[i] : nullptr; }

const char* srk_get_id(const srk_module* m) { return m ? m->id.c_str() : nullptr; }
const char* srk_get_name(const srk_module* m) { return m ? srk::kind_info(m->kind).name : nullptr; }
int srk_get_kind(const srk_module* m) { return m ? m->kind : -1; }
int srk_get_num_inputs(const srk_module* m) { return m ? (int)m->inputs.size() : -1; }
int srk_get_num_outputs(const srk_module* m) { return m ? m->n_outputs() : -1; }

int srk_get_input_label(const srk_module* m, uint8_t idx, const char** label) {
  if (!m || !label) return SRK_ERR_ARG;
  if (idx >= m->inputs.size()) return SRK_ERR_PORT;
  *label = idx < 4 ? srk::kind_info(m->kind).in_labels[idx] : nullptr;
  return SRK_OK;
}

int srk_get_output_label(const srk_module* m, uint8_t idx, const char** label) {
  if (!m || !label) return SRK_ERR_ARG;
  if ((int)idx >= m->n_outputs()) return SRK_ERR_PORT;
  *label = srk::kind_info(m->kind).out_labels[idx];
  return SRK_OK;
}

int srk_connect(srk_module* sink, uint8_t idx, srk_module* src, uint8_t port) { return guarded((sink ? sink->patch : nullptr), [&]() -> int {
  if (!sink || !src) return SRK_ERR_ARG;
  srk_patch* p = sink->patch;
  if (src->patch != p) return fail(p, SRK_ERR_ARG, "modules belong to different patches");
  if (idx >= sink->inputs.size()) return fail(p, SRK_ERR_PORT, "input index out of range");
  if ((int)port >= src->n_outputs()) return fail(p, SRK_ERR_PORT, "source port out of range");
  if (sink == src) return fail(p, SRK_ERR_SELF_LOOP, "a module cannot feed itself");
  sink->inputs[idx] = {src, port};
  touch_wiring(p);
  return SRK_OK;
}); }

int srk_disconnect(srk_module* sink, uint8_t idx) {
  if (!sink) return SRK_ERR_ARG;
  if (idx >= sink->inputs.size()) return fail(sink->patch, SRK_ERR_PORT, "input index out of range");
  sink->inputs[idx] = {nullptr, 0};
  touch_wiring(sink->patch);
  return SRK_OK;
}

int srk_disconnect_inputs(srk_module* sink) {
  if (!sink) return SRK_ERR_ARG;
  for (auto& in : sink->inputs) in = {nullptr, 0};
  touch_wiring(sink->patch);
  return SRK_OK;
}

int srk_get_input(const srk_module* sink, uint8_t idx, srk_module** src, uint8_t* port) {
  if (!sink || !src) return SRK_ERR_ARG;
  if (idx >= sink->inputs.size()) return SRK_ERR_PORT;
  *src = sink->inputs[idx].first;
  if (port) *port = sink->inputs[idx].second;
  return SRK_OK;
}

int srk_set_param_f32(srk_module* m, int pid, float value) {
  if (!m) return SRK_ERR_ARG;
  const srk::KindInfo& ki = srk::kind_info(m->kind);
  if (pid < 0 || pid >= ki.n_params) return fail(m->patch, SRK_ERR_PARAM, "unknown parameter id");
  m->param[pid] = value;
  m->param_pv[pid].clear();
  ++m->patch->param_epoch;
  if (ki.param_uniform_only[pid]) ++m->patch->table_epoch;  // baked into the program image (flags / immediates), state kept
  return SRK_OK;
}

int srk_get_param_f32(const srk_module* m, int pid, float* value) {
  if (!m || !value) return SRK_ERR_ARG;
  if (pid < 0 || pid >= srk::kind_info(m->kind).n_params) return SRK_ERR_PARAM;
  *value = m->param[pid];
  return SRK_OK;
}

int srk_set_param_f32_per_voice(srk_module* m, int pid, const float* values, size_t n) { return guarded((m ? m->patch : nullptr), [&]() -> int {
  if (!m || (!values && n)) return SRK_ERR_ARG;
  const srk::KindInfo& ki = srk::kind_info(m->kind);
  if (pid < 0 || pid >= ki.n_params) return fail(m->patch, SRK_ERR_PARAM, "unknown parameter id");
  if (ki.param_uniform_only[pid]) return fail(m->patch, SRK_ERR_PARAM, "parameter is uniform-only");
  m->param_pv[pid].assign(values, values + n);
  ++m->patch->param_epoch;
  return SRK_OK;
}); }

int srk_set_sequence(srk_module* m, const int32_t* cells, size_t n_steps) { return guarded((m ? m->patch : nullptr), [&]() -> int {
  if (!m || !cells) return SRK_ERR_ARG;
  const bool grid = m->kind == SRK_KIND_GRID_SEQUENCER, pattern = m->kind == SRK_KIND_PATTERN_SEQUENCER;
  if (!grid && !pattern) return fail(m->patch, SRK_ERR_KIND, "module has no sequence table");
  if (n_steps < 1 || n_steps > SRK_SEQ_MAX_STEPS) return fail(m->patch, SRK_ERR_ARG, "sequence length must be 1..64");
  const size_t n = (pattern ? SRK_PATTERN_ROWS : 1) * n_steps;
  for (size_t i = 0; i < n; ++i) {
    const int32_t c = cells[i];
    const bool ok = c == SRK_SEQ_NONE || (grid ? (c >= 0 && c <= 0x1FFFF) : (c == 0 || c == 1));
    if (!ok) return fail(m->patch, SRK_ERR_ARG, "bad sequence cell");
  }
  m->sequence.assign(cells, cells + n);
  m->seq_steps = n_steps;
  ++m->patch->table_epoch;
  return SRK_OK;
}); }

int srk_get_sequence(const srk_module* m, int32_t* cells, size_t cap, size_t* n_steps) {
  if (!m) return SRK_ERR_ARG;
  if (m->kind != SRK_KIND_GRID_SEQUENCER && m->kind != SRK_KIND_PATTERN_SEQUENCER) return SRK_ERR_KIND;
  if (n_steps) *n_steps = m->seq_steps;
  for (size_t i = 0; cells && i < cap && i < m->sequence.size(); ++i) cells[i] = m->sequence[i];
  return SRK_OK;
}

int srk_set_sample(srk_module* m, const float* samples, size_t n, float sample_rate) { return guarded((m ? m->patch : nullptr), [&]() -> int {
  if (!m || (!samples && n)) return SRK_ERR_ARG;
  if (m->kind != SRK_KIND_SAMPLE) return fail(m->patch, SRK_ERR_KIND, "module has no sample table");
  if (n >= (1ull << 31)) return fail(m->patch, SRK_ERR_LIMIT, "sample table longer than 2^31 - 1");
  m->wave.assign(samples, samples + n);
  m->wave_rate = sample_rate;
  m->wave_new = true;  // sample.rs:66
  ++m->patch->wave_epoch;
  ++m->patch->table_epoch;
  return SRK_OK;
}); }

int srk_load_wav(srk_module* m, const void* bytes, size_t n_bytes) { return guarded((m ? m->patch : nullptr), [&]() -> int {
  if (!m || (!bytes && n_bytes)) return SRK_ERR_ARG;
  if (m->kind != SRK_KIND_SAMPLE) return fail(m->patch, SRK_ERR_KIND, "module has no sample table");
  std::string err;
  float rate = m->wave_rate;
  const srk::WavStatus st = srk::wav_decode(bytes, n_bytes, m->wave, rate, err);
  if (st == srk::WAV_BAD_HEADER) return fail(m->patch, SRK_ERR_ARG, err.c_str());
  ++m->patch->wave_epoch;
  ++m->patch->table_epoch;
  if (st == srk::WAV_UNSUPPORTED) return fail(m->patch, SRK_ERR_UNSUPPORTED, err.c_str());  // table now empty (sample.rs:36,53)
  if (m->wave.size() >= (1ull << 31)) { m->wave.clear(); return fail(m->patch, SRK_ERR_LIMIT, "sample table longer than 2^31 - 1"); }
  m->wave_rate = rate;
  m->wave_new = true;
  return SRK_OK;
}); }

int srk_get_sample(const srk_module* m, float* samples, size_t cap, size_t* n, float* sample_rate) {
  if (!m) return SRK_ERR_ARG;
  if (m->kind != SRK_KIND_SAMPLE) return SRK_ERR_KIND;
  if (n) *n = m->wave.size();
  if (sample_rate) *sample_rate = m->wave_rate;
  for (size_t i = 0; samples && i < cap && i < m->wave.size(); ++i) samples[i] = m->wave[i];
  return SRK_OK;
}

int srk_write_wav(const char* path, const float* planar, unsigned channels, size_t n_samples, uint32_t sample_rate,
                  int bits) {
  std::string err;
  return srk::wav_write(path, planar, channels, n_samples, sample_rate, bits, err) ? SRK_OK : SRK_ERR_ARG;
}

// SynthModuleWorkspaceImpl::deserialize, ui.rs:115-134: the patch is emptied and rebuilt from the file.
int srk_patch_load_srk(srk_patch* p, const void* bytes, size_t n_bytes, size_t* n_skipped_connections) { return guarded(p, [&]() -> int {
  if (!p || (!bytes && n_bytes)) return SRK_ERR_ARG;
  srk::SrkFile f;
  std::string err;
  if (!srk::srk_file_decode(bytes, n_bytes, f, err)) return fail(p, SRK_ERR_ARG, err.c_str());
  for (const srk::SrkModule& m : f.modules)
    if (m.kind < 0) return fail(p, SRK_ERR_UNSUPPORTED, (m.variant + " is outside the hot path").c_str());
  // from here on the old patch is gone (the reference clears before it parses, ui.rs:117-122)
  p->modules.clear();
  p->owned.clear();
  p->plan.clear();
  p->cuts.clear();
  p->positions.clear();
  touch_wiring(p);
  ++p->param_epoch;
  ++p->table_epoch;
  ++p->wave_epoch;
  // unpack_modules pops from the back (ui.rs:652-660): the module list is the file's reversed
  for (size_t k = f.modules.size(); k-- > 0;) {
    const srk::SrkModule& fm = f.modules[k];
    srk_module* m = nullptr;
    int rc = srk_module_create(p, fm.kind, &m);
    if (rc != SRK_OK) return rc;
    m->id = fm.id;
    const srk::KindInfo& ki = srk::kind_info(fm.kind);
    for (int i = 0; i < ki.n_params; ++i) m->param[i] = fm.param[i];
    if (fm.kind == SRK_KIND_GRID_SEQUENCER || fm.kind == SRK_KIND_PATTERN_SEQUENCER) {
      m->sequence = fm.sequence;
      m->seq_steps = fm.seq_steps;
    }
    if (fm.kind == SRK_KIND_SAMPLE) {
      m->wave = fm.wave;
      m->wave_rate = fm.wave_rate;
      m->wave_new = false;  // a fresh voice starts rewound anyway
    }
    if (fm.has_adsr_rate) m->adsr_sample_rate = fm.adsr_sample_rate;  // set_audio_config leaves it alone (adsr.rs:69-71)
    m->init_state = fm.state;  // phase, filter memory, envelope stage, step counters, play position
  }
  // unpack_connections pops from the back too (ui.rs:673): of two entries for one input the EARLIER wins
  size_t skipped = 0;
  for (size_t k = f.connections.size(); k-- > 0;) {
    const srk::SrkConnection& c = f.connections[k];
    srk_module *sink = nullptr, *src = nullptr;
    for (srk_module* m : p->modules) {
      if (m->id == c.sink_id) sink = m;
      if (m->id == c.src_id) src = m;
    }
    if (!sink || !src || srk_connect(sink, c.sink_port, src, c.src_port) != SRK_OK) ++skipped;  // `let _ = set_input(..)`
  }
  for (const srk::SrkPosition& q : f.positions) p->positions.push_back({q.id, {q.x, q.y}});
  p->last_error.clear();
  if (n_skipped_connections) *n_skipped_connections = skipped;
  return SRK_OK;
}); }

// SynthModuleWorkspaceImpl::serialize, ui.rs:98-114.
int srk_patch_save_srk(srk_patch* p, const void** bytes, size_t* n_bytes) { return guarded(p, [&]() -> int {
  if (!p || !bytes || !n_bytes) return SRK_ERR_ARG;
  srk::SrkFile f;
  for (const srk_module* m : p->modules) {  // capture_modules: list order
    srk::SrkModule fm;
    fm.kind = m->kind;
    fm.id = m->id;
    for (int i = 0; i < srk::kMaxParams; ++i) fm.param[i] = m->param[i];
    fm.sequence = m->sequence;
    fm.seq_steps = m->seq_steps;
    fm.wave = m->wave;
    fm.wave_rate = m->wave_rate;
    fm.adsr_sample_rate = m->adsr_sample_rate;
    fm.state = m->init_state;  // a loaded file's state is written back; modules built through the ABI save new() state
    f.modules.push_back(std::move(fm));
  }
  for (const srk_module* m : p->modules)  // capture_connections: per module, inputs in index order
    for (size_t i = 0; i < m->inputs.size(); ++i)
      if (m->inputs[i].first) {
        srk::SrkConnection c;
        c.src_id = m->inputs[i].first->id;
        c.src_port = m->inputs[i].second;
        c.sink_id = m->id;
        c.sink_port = (uint8_t)i;
        f.connections.push_back(c);
      }
  for (const auto& q : p->positions)
    if (std::any_of(p->modules.begin(), p->modules.end(), [&](const srk_module* m) { return m->id == q.first; }))
      f.positions.push_back(srk::SrkPosition{q.first, q.second.first, q.second.second});
  p->saved.clear();
  srk::srk_file_encode(f, p->cfg.buffer_size, p->cfg.sample_rate, p->cfg.channels, p->saved);
  *bytes = p->saved.data();
  *n_bytes = p->saved.size();
  return SRK_OK;
}); }

int srk_plan(srk_patch* p) { return guarded(p, [&]() -> int {
  if (!p) return SRK_ERR_ARG;
  p->plan.clear();
  p->cuts.clear();
  p->planned = false;
  srk_module* output = p->find_output();
  if (!output) return fail(p, SRK_ERR_NO_OUTPUT, "patch has no Output module");  // ui.rs:76-80: empty plan
  const int n = (int)p->modules.size();
  std::vector<std::vector<int>> deps(n);
  for (int m = 0; m < n; ++m)
    for (const auto& in : p->modules[m]->inputs)
      if (in.first) deps[m].push_back(p->index_of(in.first));
  std::vector<int> order;
  std::vector<std::pair<int, int>> cuts;
  srk::plan_execution(p->index_of(output), deps, order, cuts);
  for (int i : order) p->plan.push_back(p->modules[i]);
  for (auto& c : cuts) p->cuts.emplace_back(p->modules[c.first], p->modules[c.second]);
  p->planned = true;
  return SRK_OK;
}); }

int srk_plan_get(const srk_patch* p, srk_module** out, size_t cap, size_t* n) {
  if (!p || !n) return SRK_ERR_ARG;
  *n = p->plan.size();
  if (out)
    for (size_t i = 0; i < p->plan.size() && i < cap; ++i) out[i] = p->plan[i];
  return SRK_OK;
}

int srk_plan_cuts(const srk_patch* p, srk_module** readers, srk_module** writers, size_t cap, size_t* n) {
  if (!p || !n) return SRK_ERR_ARG;
  *n = p->cuts.size();
  for (size_t i = 0; i < p->cuts.size() && i < cap; ++i) {
    if (readers) readers[i] = p->cuts[i].first;
    if (writers) writers[i] = p->cuts[i].second;
  }
  return SRK_OK;
}

int srk_set_module_order(srk_patch* p, srk_module* const* order, size_t n) { return guarded(p, [&]() -> int {
  if (!p || !order) return SRK_ERR_ARG;
  if (n != p->modules.size()) return fail(p, SRK_ERR_ARG, "order is not a permutation of the module list");
  std::vector<srk_module*> next(order, order + n), a = next, b = p->modules;
  std::sort(a.begin(), a.end());
  std::sort(b.begin(), b.end());
  if (a != b) return fail(p, SRK_ERR_ARG, "order is not a permutation of the module list");
  p->modules = next;
  touch_wiring(p);
  return SRK_OK;
}); }

int srk_render(srk_patch* p, size_t n_voices, size_t voice_offset, size_t n_samples, unsigned flags, float* stems,
               float* mix) { return guarded(p, [&]() -> int {
  if (!p) return SRK_ERR_ARG;
  return srk::engine_render(p, n_voices, voice_offset, n_samples, flags, stems, mix, nullptr, false);
}); }

int srk_render_on_stream(srk_patch* p, size_t n_voices, size_t voice_offset, size_t n_samples, unsigned flags,
                         float* stems, float* mix, void* cuda_stream) { return guarded(p, [&]() -> int {
  if (!p) return SRK_ERR_ARG;
  return srk::engine_render(p, n_voices, voice_offset, n_samples, flags, stems, mix, cuda_stream, true);
}); }

int srk_sync(srk_patch* p) { return p ? srk::engine_sync(p) : SRK_ERR_ARG; }
int srk_reset(srk_patch* p) { return p ? srk::engine_reset(p) : SRK_ERR_ARG; }

int srk_last_render_ms(srk_patch* p, float* kernel_ms, float* total_ms) {
  return p ? srk::engine_last_ms(p, kernel_ms, total_ms) : SRK_ERR_ARG;
}

uint64_t srk_launch_count(const srk_patch* p) { return p ? srk::engine_launches(p) : 0; }
uint64_t srk_state_epoch(const srk_patch* p) { return p ? srk::engine_state_epoch(p) : 0; }

int srk_get_program_info(srk_patch* p, size_t n_voices, srk_program_info* out) { return guarded(p, [&]() -> int {
  if (!p || !out) return SRK_ERR_ARG;
  return srk::engine_program_info(p, n_voices, out);
}); }

int srk_fused_source(srk_patch* p, size_t n_voices, const char** source, size_t* n_bytes) { return guarded(p, [&]() -> int {
  if (!p || !source) return SRK_ERR_ARG;
  int rc = srk::engine_fused_source(p, n_voices, p->fused_source);
  if (rc != SRK_OK) return rc;
  *source = p->fused_source.c_str();
  if (n_bytes) *n_bytes = p->fused_source.size();
  return SRK_OK;
}); }

int srk_precompile(srk_patch* p, size_t n_voices, int* compiled) { return guarded(p, [&]() -> int {
  if (!p) return SRK_ERR_ARG;
  return srk::engine_precompile(p, n_voices, compiled);
}); }

int srk_kernel_id(srk_patch* p, size_t n_voices, const char** id) { return guarded(p, [&]() -> int {
  if (!p || !id) return SRK_ERR_ARG;
  int rc = srk::engine_kernel_id(p, n_voices, p->kernel_id);
  if (rc != SRK_OK) return rc;
  *id = p->kernel_id.c_str();
  return SRK_OK;
}); }

int srk_schedule_report(srk_patch* p, const char** report) { return guarded(p, [&]() -> int {
  if (!p || !report) return SRK_ERR_ARG;
  int rc = srk::engine_tune_report(p, p->tune_report);
  if (rc != SRK_OK) return rc;
  *report = p->tune_report.c_str();
  return SRK_OK;
}); }

int srk_set_co_resident_voices(srk_patch* p, size_t n_voices) {
  if (!p) return SRK_ERR_ARG;
  p->co_resident_voices = n_voices;
  return SRK_OK;
}

int srk_state_export(srk_patch* p, const void** blob, size_t* n_bytes) { return guarded(p, [&]() -> int {
  if (!p || !blob || !n_bytes) return SRK_ERR_ARG;
  return srk::engine_state_export(p, blob, n_bytes);
}); }

int srk_state_import(srk_patch* p, const void* blob, size_t n_bytes) { return guarded(p, [&]() -> int {
  if (!p || !blob) return SRK_ERR_ARG;
  return srk::engine_state_import(p, blob, n_bytes);
}); }

int srk_get_program(srk_patch* p, size_t n_voices, srk_instr_info* instrs, size_t instr_cap, size_t* n_instr,
                    srk_wire_info* wires, size_t wire_cap, size_t* n_wires) { return guarded(p, [&]() -> int {
  if (!p) return SRK_ERR_ARG;
  return srk::engine_program_dump(p, n_voices, instrs, instr_cap, n_instr, wires, wire_cap, n_wires);
}); }

}  // extern "C"
