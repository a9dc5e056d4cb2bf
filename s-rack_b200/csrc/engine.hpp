// Host-visible interface of the device engine (engine.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

#include "patch.hpp"
#include "program.hpp"

namespace srk {

void engine_destroy(Engine* e);
int engine_render(srk_patch* patch, size_t n_voices, size_t voice_offset, size_t n_samples, unsigned flags,
                  float* stems, float* mix, void* user_stream, bool use_user_stream);
int engine_sync(srk_patch* patch);
int engine_reset(srk_patch* patch);
void engine_invalidate_state(srk_patch* patch);  // next render starts from X::new() state
int engine_last_ms(srk_patch* patch, float* kernel_ms, float* total_ms);
uint64_t engine_launches(const srk_patch* patch);
uint64_t engine_state_epoch(const srk_patch* patch);
int engine_program_info(srk_patch* patch, size_t n_voices, srk_program_info* out);
int engine_fused_source(srk_patch* patch, size_t n_voices, std::string& source);
int engine_precompile(srk_patch* patch, size_t n_voices, int* compiled);
int engine_tune_report(srk_patch* patch, std::string& report);
int engine_kernel_id(srk_patch* patch, size_t n_voices, std::string& id);
int engine_state_export(srk_patch* patch, const void** blob, size_t* n_bytes);
int engine_state_import(srk_patch* patch, const void* blob, size_t n_bytes);
int engine_program_dump(srk_patch* patch, size_t n_voices, srk_instr_info* instrs, size_t instr_cap, size_t* n_instr,
                        srk_wire_info* wires, size_t wire_cap, size_t* n_wires);

}  // namespace srk
