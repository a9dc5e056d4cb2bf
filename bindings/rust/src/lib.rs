//! Rust host side of the B200 voice renderer: a thin `extern "C"` layer over
//! `include/srack_b200.h` plus a facade whose names follow the reference's
//! `synth` module (`src/synth.rs`): `AudioConfig`, `Module` (the `SynthModule`
//! trait's wiring/introspection methods, `synth.rs:222-263`), `Patch::plan`
//! (`plan_execution`, `synth.rs:128`) and `Patch::execute` (`execute`,
//! `synth.rs:97`, for `n_voices` instances at once).
//!
//! Source only: this image has no rustc/cargo, so this crate is exercised by
//! review, not by CI.  Every call below maps 1:1 to a C entry point that *is*
//! tested (tests/test_abi.py, tests/test_gpu_parity.py through ctypes).
#![allow(non_camel_case_types)]

use std::ffi::{c_char, c_int, c_uint, c_void, CStr, CString};
use std::ptr;

#[repr(C)]
pub struct srk_patch { _p: [u8; 0] }
#[repr(C)]
pub struct srk_module { _p: [u8; 0] }

/// `src/synth.rs:20-25`
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct AudioConfig {
    pub sample_rate: u16,
    pub buffer_size: usize,
    pub channels: u8,
}

pub const SRK_RENDER_DEVICE_OUT: c_uint = 1 << 0;
pub const SRK_RENDER_ASYNC: c_uint = 1 << 1;

extern "C" {
    fn srk_status_string(status: c_int) -> *const c_char;
    fn srk_patch_create(cfg: *const AudioConfig, out: *mut *mut srk_patch) -> c_int;
    fn srk_patch_destroy(patch: *mut srk_patch);
    fn srk_set_audio_config(patch: *mut srk_patch, cfg: *const AudioConfig) -> c_int;
    fn srk_set_seed(patch: *mut srk_patch, seed: u64) -> c_int;
    fn srk_set_device(patch: *mut srk_patch, device: c_int) -> c_int;
    fn srk_last_error(patch: *const srk_patch) -> *const c_char;
    fn srk_module_create_by_name(patch: *mut srk_patch, name: *const c_char, out: *mut *mut srk_module) -> c_int;
    fn srk_module_remove(patch: *mut srk_patch, module: *mut srk_module) -> c_int;
    fn srk_get_id(m: *const srk_module) -> *const c_char;
    fn srk_get_name(m: *const srk_module) -> *const c_char;
    fn srk_get_num_inputs(m: *const srk_module) -> c_int;
    fn srk_get_num_outputs(m: *const srk_module) -> c_int;
    fn srk_get_input_label(m: *const srk_module, idx: u8, label: *mut *const c_char) -> c_int;
    fn srk_get_output_label(m: *const srk_module, idx: u8, label: *mut *const c_char) -> c_int;
    fn srk_connect(sink: *mut srk_module, input_idx: u8, src: *mut srk_module, src_port: u8) -> c_int;
    fn srk_disconnect(sink: *mut srk_module, input_idx: u8) -> c_int;
    fn srk_disconnect_inputs(sink: *mut srk_module) -> c_int;
    fn srk_get_input(sink: *const srk_module, input_idx: u8, src: *mut *mut srk_module, port: *mut u8) -> c_int;
    fn srk_set_param_f32(m: *mut srk_module, param_id: c_int, value: f32) -> c_int;
    fn srk_set_param_f32_per_voice(m: *mut srk_module, param_id: c_int, values: *const f32, n: usize) -> c_int;
    fn srk_set_sequence(m: *mut srk_module, cells: *const i32, n_steps: usize) -> c_int;
    fn srk_load_wav(m: *mut srk_module, wav_bytes: *const c_void, n_bytes: usize) -> c_int;
    fn srk_set_sample(m: *mut srk_module, samples: *const f32, n: usize, sample_rate: f32) -> c_int;
    fn srk_write_wav(path: *const c_char, planar: *const f32, channels: c_uint, n_samples: usize,
                     sample_rate: u32, bits: c_int) -> c_int;
    fn srk_patch_load_srk(patch: *mut srk_patch, bytes: *const c_void, n_bytes: usize, n_skipped: *mut usize) -> c_int;
    fn srk_patch_save_srk(patch: *mut srk_patch, bytes: *mut *const c_void, n_bytes: *mut usize) -> c_int;
    fn srk_state_epoch(patch: *const srk_patch) -> u64;
    fn srk_state_export(patch: *mut srk_patch, blob: *mut *const c_void, n_bytes: *mut usize) -> c_int;
    fn srk_state_import(patch: *mut srk_patch, blob: *const c_void, n_bytes: usize) -> c_int;
    fn srk_set_co_resident_voices(patch: *mut srk_patch, n_voices: usize) -> c_int;
    fn srk_plan(patch: *mut srk_patch) -> c_int;
    fn srk_plan_get(patch: *const srk_patch, out: *mut *mut srk_module, cap: usize, n: *mut usize) -> c_int;
    fn srk_render(patch: *mut srk_patch, n_voices: usize, voice_offset: usize, n_samples: usize,
                  flags: c_uint, stems: *mut f32, mix: *mut f32) -> c_int;
    fn srk_render_on_stream(patch: *mut srk_patch, n_voices: usize, voice_offset: usize, n_samples: usize,
                            flags: c_uint, stems: *mut f32, mix: *mut f32, stream: *mut c_void) -> c_int;
    fn srk_sync(patch: *mut srk_patch) -> c_int;
    fn srk_reset(patch: *mut srk_patch) -> c_int;
}

/// The reference's `Err(())` carries nothing; here it carries the status and text.
#[derive(Debug)]
pub struct Error { pub status: i32, pub detail: String }

fn cstr(p: *const c_char) -> String {
    if p.is_null() { String::new() } else { unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned() }
}

/// One module of a patch (`SharedSynthModule`, `synth.rs:270`); identity is the
/// handle, like `shared_are_eq` (`synth.rs:272`).  Owned by its `Patch`.
#[derive(Clone, Copy, PartialEq, Eq, Hash)]
pub struct Module { h: *mut srk_module, patch: *mut srk_patch }

impl Module {
    fn check(&self, rc: c_int) -> Result<(), Error> {
        if rc == 0 { Ok(()) } else {
            Err(Error { status: rc, detail: cstr(unsafe { srk_last_error(self.patch) }) })
        }
    }
    pub fn get_id(&self) -> String { cstr(unsafe { srk_get_id(self.h) }) }
    pub fn get_name(&self) -> String { cstr(unsafe { srk_get_name(self.h) }) }
    pub fn get_num_inputs(&self) -> u8 { unsafe { srk_get_num_inputs(self.h) as u8 } }
    pub fn get_num_outputs(&self) -> u8 { unsafe { srk_get_num_outputs(self.h) as u8 } }
    pub fn get_input_label(&self, idx: u8) -> Result<Option<String>, Error> {
        let mut p = ptr::null();
        self.check(unsafe { srk_get_input_label(self.h, idx, &mut p) })?;
        Ok(if p.is_null() { None } else { Some(cstr(p)) })
    }
    pub fn get_output_label(&self, idx: u8) -> Result<Option<String>, Error> {
        let mut p = ptr::null();
        self.check(unsafe { srk_get_output_label(self.h, idx, &mut p) })?;
        Ok(if p.is_null() { None } else { Some(cstr(p)) })
    }
    /// `SynthModule::set_input`, `synth.rs:234-239`
    pub fn set_input(&self, input_idx: u8, src: &Module, src_port: u8) -> Result<(), Error> {
        self.check(unsafe { srk_connect(self.h, input_idx, src.h, src_port) })
    }
    pub fn disconnect_input(&self, input_idx: u8) -> Result<(), Error> {
        self.check(unsafe { srk_disconnect(self.h, input_idx) })
    }
    pub fn disconnect_inputs(&self) { unsafe { srk_disconnect_inputs(self.h); } }
    /// `SynthModule::get_input`, `synth.rs:228`
    pub fn get_input(&self, input_idx: u8) -> Result<Option<(Module, u8)>, Error> {
        let (mut src, mut port) = (ptr::null_mut(), 0u8);
        self.check(unsafe { srk_get_input(self.h, input_idx, &mut src, &mut port) })?;
        Ok(if src.is_null() { None } else { Some((Module { h: src, patch: self.patch }, port)) })
    }
    /// Struct fields the reference mutates from `ui()` (ids: `enum srk_param`).
    pub fn set_param(&self, param_id: i32, value: f32) -> Result<(), Error> {
        self.check(unsafe { srk_set_param_f32(self.h, param_id, value) })
    }
    /// A sequencer's step table (`sequencer.rs:18,341`): `n_steps` cells (grid) or 8 rows x `n_steps` (pattern).
    pub fn set_sequence(&self, cells: &[i32], n_steps: usize) -> Result<(), Error> {
        self.check(unsafe { srk_set_sequence(self.h, cells.as_ptr(), n_steps) })
    }
    /// `WaveBox::load` (`sample.rs:32-69`): the bytes of a WAV file into a Sample module's table.
    pub fn load_wav(&self, wav: &[u8]) -> Result<(), Error> {
        self.check(unsafe { srk_load_wav(self.h, wav.as_ptr() as *const c_void, wav.len()) })
    }
    /// The decoded `WaveBox` directly (`sample.rs:16-20`).
    pub fn set_sample(&self, samples: &[f32], sample_rate: f32) -> Result<(), Error> {
        self.check(unsafe { srk_set_sample(self.h, samples.as_ptr(), samples.len(), sample_rate) })
    }
    pub fn set_param_per_voice(&self, param_id: i32, values: &[f32]) -> Result<(), Error> {
        self.check(unsafe { srk_set_param_f32_per_voice(self.h, param_id, values.as_ptr(), values.len()) })
    }
}

/// `get_inputs`, `synth.rs:214-218`
pub fn get_inputs(m: &Module) -> Vec<Option<(Module, u8)>> {
    (0..m.get_num_inputs()).map(|i| m.get_input(i).unwrap()).collect()
}

/// The module list + plan (`SynthModuleWorkspaceImpl`, `ui.rs:51-60`) and the device engine.
pub struct Patch { h: *mut srk_patch, cfg: AudioConfig }

// A patch is single-threaded (the caller serialises, as `Mutex<plan>` does in main.rs:60).
unsafe impl Send for Patch {}

impl Patch {
    pub fn new(cfg: &AudioConfig) -> Result<Patch, Error> {
        let mut h = ptr::null_mut();
        let rc = unsafe { srk_patch_create(cfg, &mut h) };
        if rc != 0 { return Err(Error { status: rc, detail: cstr(unsafe { srk_status_string(rc) }) }); }
        Ok(Patch { h, cfg: *cfg })
    }
    fn check(&self, rc: c_int) -> Result<(), Error> {
        if rc == 0 { Ok(()) } else { Err(Error { status: rc, detail: cstr(unsafe { srk_last_error(self.h) }) }) }
    }
    pub fn audio_config(&self) -> AudioConfig { self.cfg }
    pub fn set_audio_config(&mut self, cfg: &AudioConfig) -> Result<(), Error> {
        self.cfg = *cfg;
        self.check(unsafe { srk_set_audio_config(self.h, cfg) })
    }
    pub fn set_seed(&mut self, seed: u64) -> Result<(), Error> { self.check(unsafe { srk_set_seed(self.h, seed) }) }
    pub fn set_device(&mut self, device: i32) -> Result<(), Error> { self.check(unsafe { srk_set_device(self.h, device) }) }
    /// A catalog closure (`get_catalog`, `synth.rs:421-515`) or `OutputModule::new` (`main.rs:130`).
    pub fn add(&mut self, catalog_name: &str) -> Result<Module, Error> {
        let name = CString::new(catalog_name).unwrap();
        let mut m = ptr::null_mut();
        self.check(unsafe { srk_module_create_by_name(self.h, name.as_ptr(), &mut m) })?;
        Ok(Module { h: m, patch: self.h })
    }
    pub fn remove(&mut self, m: Module) -> Result<(), Error> { self.check(unsafe { srk_module_remove(self.h, m.h) }) }
    /// `plan_execution(output, &all_modules, &mut plan)`, `synth.rs:128-212`
    pub fn plan(&mut self) -> Result<Vec<Module>, Error> {
        self.check(unsafe { srk_plan(self.h) })?;
        let mut n = 0usize;
        self.check(unsafe { srk_plan_get(self.h, ptr::null_mut(), 0, &mut n) })?;
        let mut raw = vec![ptr::null_mut(); n];
        self.check(unsafe { srk_plan_get(self.h, raw.as_mut_ptr(), n, &mut n) })?;
        Ok(raw.into_iter().map(|h| Module { h, patch: self.h }).collect())
    }
    /// `execute(&plan)` (`synth.rs:97-101`) for `n_samples / buffer_size` blocks of `n_voices`
    /// instances, then what the audio callback reads from `OutputModule.bufs` (`main.rs:64-75`):
    /// `mix` is `[channels][n_samples]`, `stems` (optional) `[channels][n_samples][n_voices]`.
    pub fn execute(&mut self, n_voices: usize, voice_offset: usize, n_samples: usize,
                   stems: Option<&mut [f32]>, mix: Option<&mut [f32]>) -> Result<(), Error> {
        let c = self.cfg.channels as usize;
        if let Some(s) = &stems { assert!(s.len() >= c * n_samples * n_voices); }
        if let Some(m) = &mix { assert!(m.len() >= c * n_samples); }
        let sp = stems.map_or(ptr::null_mut(), |s| s.as_mut_ptr());
        let mp = mix.map_or(ptr::null_mut(), |m| m.as_mut_ptr());
        self.check(unsafe { srk_render(self.h, n_voices, voice_offset, n_samples, 0, sp, mp) })
    }
    /// Device-pointer variant on a caller-owned CUDA stream (e.g. feeding NCCL directly).
    /// # Safety
    /// `stems`/`mix` must be device pointers of sufficient size on this patch's device.
    pub unsafe fn execute_device(&mut self, n_voices: usize, voice_offset: usize, n_samples: usize,
                                 stems: *mut f32, mix: *mut f32, stream: *mut c_void) -> Result<(), Error> {
        self.check(srk_render_on_stream(self.h, n_voices, voice_offset, n_samples,
                                        SRK_RENDER_DEVICE_OUT | SRK_RENDER_ASYNC, stems, mix, stream))
    }
    pub fn sync(&mut self) -> Result<(), Error> { self.check(unsafe { srk_sync(self.h) }) }
    pub fn reset(&mut self) -> Result<(), Error> { self.check(unsafe { srk_reset(self.h) }) }
    /// `SynthModuleWorkspaceImpl::deserialize` (`ui.rs:115-134`): replace the patch by a `.srk` file's.
    /// Returns how many connections were skipped (unknown ids / bad ports).  Call `plan()` afterwards.
    pub fn load_srk(&mut self, bytes: &[u8]) -> Result<usize, Error> {
        let mut skipped = 0usize;
        self.check(unsafe { srk_patch_load_srk(self.h, bytes.as_ptr() as *const c_void, bytes.len(), &mut skipped) })?;
        Ok(skipped)
    }
    /// `SynthModuleWorkspaceImpl::serialize` (`ui.rs:98-114`).
    pub fn save_srk(&mut self) -> Result<Vec<u8>, Error> {
        let (mut p, mut n) = (ptr::null(), 0usize);
        self.check(unsafe { srk_patch_save_srk(self.h, &mut p, &mut n) })?;
        Ok(unsafe { std::slice::from_raw_parts(p as *const u8, n) }.to_vec())
    }

    /// Checkpoint of a render in progress: every voice's DSP state (what the reference's modules serialize with
    /// `#[derive(Serialize)]`, `ui.rs:98-114`), the history of cut wires and the sample index, for the voice range
    /// last rendered.
    pub fn state_export(&mut self) -> Result<Vec<u8>, Error> {
        let (mut p, mut n) = (ptr::null(), 0usize);
        self.check(unsafe { srk_state_export(self.h, &mut p, &mut n) })?;
        Ok(unsafe { std::slice::from_raw_parts(p as *const u8, n) }.to_vec())
    }

    /// Resume: the next `execute` of the blob's voice range continues bit for bit (`ui.rs:115-134`).
    pub fn state_import(&mut self, blob: &[u8]) -> Result<(), Error> {
        self.check(unsafe { srk_state_import(self.h, blob.as_ptr() as *const c_void, blob.len()) })
    }

    /// How often the voice state was (re)initialised -- implicitly too, by an `execute` of another voice range.
    pub fn state_epoch(&self) -> u64 {
        unsafe { srk_state_epoch(self.h) }
    }

    /// Voices other patches render on this device at the same time (the launch is scheduled for the sum).
    pub fn set_co_resident_voices(&mut self, n_voices: usize) -> Result<(), Error> {
        self.check(unsafe { srk_set_co_resident_voices(self.h, n_voices) })
    }
}

/// WAV export of a render: `planar` is `[channels][n_samples]` as `Patch::execute` writes `mix`.
pub fn write_wav(path: &str, planar: &[f32], channels: u32, sample_rate: u32, bits: i32) -> Result<(), Error> {
    let c = std::ffi::CString::new(path).map_err(|_| Error { status: 1, detail: "path contains NUL".into() })?;
    let n = planar.len() / channels.max(1) as usize;
    let rc = unsafe { srk_write_wav(c.as_ptr(), planar.as_ptr(), channels, n, sample_rate, bits) };
    if rc == 0 { Ok(()) } else { Err(Error { status: rc, detail: "cannot write WAV".into() }) }
}

impl Drop for Patch {
    fn drop(&mut self) { unsafe { srk_patch_destroy(self.h) } }
}
