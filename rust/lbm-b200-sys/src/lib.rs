//! Raw bindings to `include/lbm_b200.h` — the CUDA replacement of the wgpu objects owned by
//! `simuverse::fluid::D2Q9Node` (simuverse/src/fluid/d2q9_node.rs:13-27).
//!
//! The `#[repr(C)]` structs are the reference's own Pod types (`LbmUniform`, `LatticeInfo`,
//! `FieldUniform`, `ParticleUniform`, `TrajectoryParticle`), so inside simuverse these can simply be
//! `pub use crate::{fluid::{LbmUniform, LatticeInfo}, FieldUniform, ...}` instead of redefinitions.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)]
pub struct LbmSim {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Copy, Clone, Debug, Default)]
pub struct LbmDesc {
    pub struct_size: u32,
    pub nx: i32,
    pub ny: i32,
    pub lattice_pixel_size: i32,
    pub canvas_w: i32,
    pub canvas_h: i32,
    pub device: i32,
    pub rank: i32,
    pub world: i32,
    pub flags: u32,
    pub max_particles: i32,
}

#[repr(C)]
#[derive(Copy, Clone)]
pub struct LbmIpcBlob {
    pub bytes: [u8; 256],
}

#[repr(C)]
#[derive(Copy, Clone, Debug, bytemuck::Pod, bytemuck::Zeroable)]
pub struct LbmUniform {
    pub tau: f32,
    pub omega: f32,
    pub fluid_ty: i32,
    pub soa_offset: i32,
    pub e_w_max: [[f32; 4]; 9],
    pub inversed_direction: [[i32; 4]; 9],
}

#[repr(C)]
#[derive(Copy, Clone, Debug, bytemuck::Pod, bytemuck::Zeroable)]
pub struct LatticeInfo {
    pub material: i32,
    pub block_iter: i32,
    pub vx: f32,
    pub vy: f32,
}

#[repr(C)]
#[derive(Copy, Clone, Debug, bytemuck::Pod, bytemuck::Zeroable)]
pub struct FieldUniform {
    pub lattice_size: [i32; 2],
    pub lattice_pixel_size: [f32; 2],
    pub canvas_size: [i32; 2],
    pub proj_ratio: [f32; 2],
    pub ndc_pixel: [f32; 2],
    pub speed_ty: i32,
    pub _padding: f32,
}

#[repr(C)]
#[derive(Copy, Clone, bytemuck::Pod, bytemuck::Zeroable)]
pub struct ParticleUniform {
    pub color: [f32; 4],
    pub num: [i32; 2],
    pub point_size: i32,
    pub life_time: f32,
    pub fade_out_factor: f32,
    pub speed_factor: f32,
    pub color_ty: i32,
    pub is_only_update_pos: i32,
}

#[repr(C)]
#[derive(Copy, Clone, bytemuck::Pod, bytemuck::Zeroable)]
pub struct TrajectoryParticle {
    pub pos: [f32; 2],
    pub pos_initial: [f32; 2],
    pub life_time: f32,
    pub fade: f32,
}

#[repr(C)]
#[derive(Copy, Clone, bytemuck::Pod, bytemuck::Zeroable)]
pub struct Pixel {
    pub alpha: f32,
    pub velocity_x: f32,
    pub velocity_y: f32,
}

pub const LBM_OK: c_int = 0;
pub const LBM_FLAG_MACRO_EVERY_STEP: u32 = 0x1;
pub const LBM_FLAG_KERNEL_GENERIC: u32 = 0x2;
pub const LBM_FLAG_NO_GRAPH: u32 = 0x4;
pub const LBM_FLAG_AA: u32 = 0x8;
pub const LBM_FLAG_NO_FUSE: u32 = 0x10;
pub const LBM_MACRO_F32_PLANES: i32 = 0;
pub const LBM_MACRO_RGBA16F: i32 = 1;

unsafe extern "C" {
    pub fn lbm_abi_version() -> c_int;
    pub fn lbm_device_count() -> c_int;
    pub fn lbm_create(desc: *const LbmDesc, out: *mut *mut LbmSim) -> c_int;
    pub fn lbm_destroy(sim: *mut LbmSim);
    pub fn lbm_last_error(sim: *const LbmSim) -> *const c_char;
    pub fn lbm_status_string(status: c_int) -> *const c_char;

    pub fn lbm_write_uniform(sim: *mut LbmSim, u: *const LbmUniform) -> c_int;
    pub fn lbm_write_field_uniform(sim: *mut LbmSim, f: *const FieldUniform) -> c_int;
    pub fn lbm_write_lattice_info(sim: *mut LbmSim, byte_offset: u64, src: *const c_void, nbytes: u64) -> c_int;
    pub fn lbm_generate_lattice_info(sim: *mut LbmSim, kind: i32, seed: u64, solid_fraction: f32) -> c_int;

    pub fn lbm_reset(sim: *mut LbmSim) -> c_int;
    pub fn lbm_step(sim: *mut LbmSim, swap_index: i32) -> c_int;
    pub fn lbm_step_n(sim: *mut LbmSim, n: i32) -> c_int;
    pub fn lbm_compute_frames(sim: *mut LbmSim, n_frames: i32) -> c_int;
    pub fn lbm_swap_index(sim: *const LbmSim) -> c_int;
    pub fn lbm_sync(sim: *mut LbmSim) -> c_int;

    pub fn lbm_slab_rows(sim: *const LbmSim, y0: *mut i32, rows: *mut i32) -> c_int;
    pub fn lbm_read_distributions(sim: *mut LbmSim, which: i32, dst: *mut f32) -> c_int;
    pub fn lbm_write_distributions(sim: *mut LbmSim, which: i32, src: *const f32) -> c_int;
    pub fn lbm_read_macro(sim: *mut LbmSim, format: i32, dst: *mut c_void) -> c_int;
    pub fn lbm_read_macro_async(sim: *mut LbmSim, dst: *mut c_void) -> c_int;
    pub fn lbm_read_curl(sim: *mut LbmSim, dst: *mut c_void) -> c_int;
    pub fn lbm_read_present(sim: *mut LbmSim, row0: i32, rows: i32, dst: *mut f32) -> c_int;
    pub fn lbm_read_lattice_info(sim: *mut LbmSim, dst: *mut LatticeInfo) -> c_int;
    pub fn lbm_total_mass(sim: *mut LbmSim, which: i32, out: *mut f64) -> c_int;

    pub fn lbm_write_particle_uniform(sim: *mut LbmSim, pu: *const ParticleUniform) -> c_int;
    pub fn lbm_particles_write(sim: *mut LbmSim, src: *const TrajectoryParticle, count: u64) -> c_int;
    pub fn lbm_particles_update(sim: *mut LbmSim) -> c_int;
    pub fn lbm_particles_read(sim: *mut LbmSim, dst: *mut TrajectoryParticle, count: u64) -> c_int;
    pub fn lbm_canvas_clear(sim: *mut LbmSim) -> c_int;
    pub fn lbm_canvas_fade(sim: *mut LbmSim) -> c_int;
    pub fn lbm_canvas_read(sim: *mut LbmSim, dst: *mut Pixel) -> c_int;

    pub fn lbm_ipc_export(sim: *mut LbmSim, out: *mut LbmIpcBlob) -> c_int;
    pub fn lbm_ipc_attach(sim: *mut LbmSim, up: *const LbmIpcBlob, down: *const LbmIpcBlob) -> c_int;

    pub fn lbm_refresh_previous(sim: *mut LbmSim) -> i32;
    pub fn lbm_sweep_blocks(h: i32, rows_per_block: i32, out: *mut i32, cap: i32, n_edge: *mut i32) -> i32;
    pub fn lbm_scan_lattice_info_write(nx: i32, ny: i32, byte_offset: u64, src: *const c_void, nbytes: u64, armed: *mut i32, border_solid: *mut i32) -> i32;
    pub fn lbm_sweep_blocks_tail(h: i32, rows_per_block: i32, tail_rows: i32, tail_rows_per_block: i32, out: *mut i32, cap: i32, n_edge: *mut i32) -> i32;
    pub fn lbm_launch_count(sim: *const LbmSim) -> u64;
    pub fn lbm_fused_sweep_count(sim: *const LbmSim) -> u64;
    pub fn lbm_sweep_uses_masked_path(sim: *const LbmSim) -> c_int;
    pub fn lbm_last_step_n_ms(sim: *mut LbmSim, ms: *mut f32) -> c_int;
    pub fn lbm_edge_wait_stats(sim: *mut LbmSim, total_ns: *mut u64, n_waits: *mut u64) -> c_int;
    pub fn lbm_stream(sim: *mut LbmSim) -> *mut c_void;
}
