//! `CudaD2Q9Node` — drop-in for `simuverse::fluid::D2Q9Node` (simuverse/src/fluid/d2q9_node.rs:13-313) on top of
//! `lbm-b200-sys`.
//!
//! File placement: `simuverse/src/fluid/cuda_d2q9_node.rs`, declared in `fluid/mod.rs` next to `mod d2q9_node;`
//! (`crate::`/`super::` paths below are the reference's own).  Every wgpu object the reference node owns — the two
//! ping-pong storage buffers, `info_buf`, both uniform buffers, `macro_tex`, the init / collide_stream / boundary
//! pipelines (d2q9_node.rs:13-27) — is replaced by ONE `LbmSim` handle; each call site that touched them maps to one C
//! call (table at the top of include/lbm_b200.h).  What the node still does on the host is what the reference does on
//! the host: the CPU mirror `lattice_info_data`, the obstacle patch and the force-cell sampling.
//!
//! UNBUILT in this repository (no cargo / rustc in the image, and the simuverse crate needs ~400 dependencies);
//! the same logic is executed and tested through the C++ and Python mirrors (include/d2q9_node.hpp,
//! simuverse_b200/d2q9_node.py), which call the same C ABI.
use alloc::{string::String, vec::Vec};
use core::ffi::{CStr, c_void};

use lbm_b200_sys as sys;

use super::{LatticeInfo, LatticeType, OBSTACLE_RADIUS, init_lattice_material, is_sd_sphere};
use crate::{FieldAnimationType, FieldUniform, SettingObj, fluid::LbmUniform};

/// `Err(message)` instead of the reference's panics (util/shader.rs:81,123): nothing unwinds across the FFI.
pub type LbmResult<T> = Result<T, String>;

pub struct CudaD2Q9Node {
    sim: *mut sys::LbmSim,
    pub lattice: wgpu::Extent3d,
    pub lattice_pixel_size: u32,
    animation_ty: FieldAnimationType,
    pub lbm_uniform_data: LbmUniform,
    pub field_uniform_data: FieldUniform,
    /// CPU mirror of the info buffer; like the reference's, it is NOT updated by force-cell writes nor by the
    /// device-side block_iter countdown (d2q9_node.rs:20, collide_stream.wgsl:55-62).
    pub lattice_info_data: Vec<LatticeInfo>,
    pub workgroup_count: (u32, u32, u32),
}

// The handle is used from the winit thread only, like the reference's node (app_handler.rs:34).
impl CudaD2Q9Node {
    fn check(&self, status: i32) -> LbmResult<()> {
        if status == sys::LBM_OK {
            return Ok(());
        }
        // SAFETY: both calls return NUL-terminated strings owned by the library / the handle.
        let (what, detail) = unsafe {
            (
                CStr::from_ptr(sys::lbm_status_string(status)).to_string_lossy().into_owned(),
                CStr::from_ptr(sys::lbm_last_error(self.sim)).to_string_lossy().into_owned(),
            )
        };
        Err(alloc::format!("lbm_b200 status {status}: {what}: {detail}"))
    }

    /// `D2Q9Node::new` (d2q9_node.rs:31-209).  `max_particles` > 0 makes the handle own the tracer-particle and
    /// canvas buffers the reference borrows from `SettingObj` / the app (fluid_simulator.rs:30,99-101) and turns on
    /// the per-update macro texture the particles sample.
    pub fn new(
        app: &app_surface::AppSurface,
        canvas_size: glam::UVec2,
        setting: &SettingObj,
        max_particles: i32,
    ) -> LbmResult<Self> {
        let lattice_pixel_size = (2.0 * app.scale_factor).ceil() as u32;
        let lattice = wgpu::Extent3d {
            width: canvas_size.x / lattice_pixel_size,
            height: canvas_size.y / lattice_pixel_size,
            depth_or_array_layers: 1,
        };
        let desc = sys::LbmDesc {
            struct_size: core::mem::size_of::<sys::LbmDesc>() as u32,
            nx: lattice.width as i32,
            ny: lattice.height as i32,
            lattice_pixel_size: lattice_pixel_size as i32,
            canvas_w: canvas_size.x as i32,
            canvas_h: canvas_size.y as i32,
            device: -1,
            rank: 0,
            world: 1,
            flags: if max_particles > 0 { sys::LBM_FLAG_MACRO_EVERY_STEP } else { 0 },
            max_particles,
        };
        let mut sim = core::ptr::null_mut();
        // SAFETY: desc is a fully initialised LbmDesc; `sim` receives the handle or stays null.
        let status = unsafe { sys::lbm_create(&desc, &mut sim) };
        if status != sys::LBM_OK {
            let detail = unsafe { CStr::from_ptr(sys::lbm_last_error(core::ptr::null())) };
            return Err(alloc::format!("lbm_create: status {status}: {}", detail.to_string_lossy()));
        }

        let fluid_ty = (setting.animation_type == FieldAnimationType::LidDrivenCavity) as i32;
        let tau = 3.0 * setting.fluid_viscosity + 0.5; // d2q9_node.rs:50
        let lbm_uniform_data = LbmUniform::new(tau, fluid_ty, (lattice.width * lattice.height) as i32);
        // d2q9_node.rs:61-64 (no LBM shader reads proj_ratio / ndc_pixel, but they are part of the reference's 48 bytes)
        let (_, sx, sy) = crate::util::matrix_helper::fullscreen_factor(
            (canvas_size.x as f32, canvas_size.y as f32).into(),
            75.0 / 180.0 * core::f32::consts::PI,
        );
        let field_uniform_data = FieldUniform {
            lattice_size: [lattice.width as i32, lattice.height as i32],
            lattice_pixel_size: [lattice_pixel_size as f32; 2],
            canvas_size: [canvas_size.x as i32, canvas_size.y as i32],
            proj_ratio: [sx, sy],
            ndc_pixel: [sx * 2.0 / canvas_size.x as f32, sy * 2.0 / canvas_size.y as f32],
            speed_ty: 1,
            _padding: 0.0,
        };
        let mut node = CudaD2Q9Node {
            sim,
            lattice,
            lattice_pixel_size,
            animation_ty: setting.animation_type,
            lbm_uniform_data,
            field_uniform_data,
            lattice_info_data: init_lattice_material(lattice, setting.animation_type),
            workgroup_count: (lattice.width.div_ceil(64), lattice.height.div_ceil(4), 1),
        };
        // LbmUniform / FieldUniform / LatticeInfo are the byte images the C ABI expects (include/lbm_wire.h).
        node.check(unsafe { sys::lbm_write_uniform(node.sim, (&node.lbm_uniform_data as *const LbmUniform).cast()) })?;
        node.check(unsafe {
            sys::lbm_write_field_uniform(node.sim, (&node.field_uniform_data as *const FieldUniform).cast())
        })?;
        node.upload_info(0, &node.lattice_info_data)?;
        node.reset()?; // d2q9_node.rs:206
        Ok(node)
    }

    /// `queue.write_buffer(&info_buf, offset, cells)` (d2q9_node.rs:244,250-254,298).  Returns as soon as the bytes
    /// are staged: the library copies them and orders the upload before the next step.
    fn upload_info(&self, byte_offset: u64, cells: &[LatticeInfo]) -> LbmResult<()> {
        let bytes: &[u8] = bytemuck::cast_slice(cells);
        self.check(unsafe {
            sys::lbm_write_lattice_info(self.sim, byte_offset, bytes.as_ptr() as *const c_void, bytes.len() as u64)
        })
    }

    /// `D2Q9Node::reset` (d2q9_node.rs:211-213): init.wgsl.
    pub fn reset(&mut self) -> LbmResult<()> {
        self.check(unsafe { sys::lbm_reset(self.sim) })
    }

    /// `add_obstacle` (d2q9_node.rs:215-245): a disc of radius 28 around the centre of cell (x, y) is stamped into
    /// the mirror and the 56 full rows it spans are uploaded again.
    pub fn add_obstacle(&mut self, x: u32, y: u32) -> LbmResult<()> {
        let disc = LatticeInfo { material: LatticeType::Obstacle as i32, block_iter: -1, vx: 0.0, vy: 0.0 };
        let radius = OBSTACLE_RADIUS as u32;
        let (first_row, last_row) = (y - radius, y + radius); // rows first_row .. last_row (exclusive)
        let width = self.lattice.width;
        let centre = glam::Vec2::new(x as f32 + 0.5, y as f32 + 0.5);
        let (first, last) = ((width * first_row) as usize, (width * last_row) as usize);
        for index in first..last {
            let (cx, cy) = (index as u32 % width, index as u32 / width);
            let from_centre = glam::Vec2::new(cx as f32 + 0.5, cy as f32 + 0.5) - centre;
            if is_sd_sphere(&from_centre, OBSTACLE_RADIUS) {
                self.lattice_info_data[index] = disc;
            }
        }
        self.upload_info(first as u64 * 16, &self.lattice_info_data[first..last])
    }

    /// `reset_lattice_info` (d2q9_node.rs:247-261): the Poiseuille preset forgets painted obstacles, then init.
    pub fn reset_lattice_info(&mut self) -> LbmResult<()> {
        if self.animation_ty == FieldAnimationType::Poiseuille {
            self.lattice_info_data = init_lattice_material(self.lattice, self.animation_ty);
            self.upload_info(0, &self.lattice_info_data)?;
        }
        self.reset()
    }

    /// `add_external_force` (d2q9_node.rs:263-300): force cells (material 6, 90 updates to live) along the segment
    /// pre_pos -> pos, one 16-byte write per sample point; the mirror is left alone, as in the reference.
    pub fn add_external_force(&mut self, pos: glam::Vec2, pre_pos: glam::Vec2) -> LbmResult<()> {
        let travelled = pos.distance(pre_pos);
        let strength = (0.1 * (travelled / 20.0)).min(0.12);
        let heading = (pos.y - pre_pos.y).atan2(pos.x - pre_pos.x); // [-pi, pi]
        let (sin, cos) = (heading.sin(), heading.cos());
        let cell = [LatticeInfo {
            material: LatticeType::ExternalForce as i32,
            block_iter: 90,
            vx: strength * cos,
            vy: strength * sin,
        }];
        let samples = (travelled / (self.lattice_pixel_size - 1) as f32).ceil();
        let stride = travelled / samples;
        let (width, height) = (self.lattice.width, self.lattice.height);
        for i in 0..samples as i32 {
            let along = stride * i as f32;
            let p = glam::Vec2::new(pre_pos.x + along * cos, pre_pos.y + along * sin).round();
            let (cx, cy) = (p.x as u32 / self.lattice_pixel_size, p.y as u32 / self.lattice_pixel_size);
            let inside = (1..width - 2).contains(&cx) && (1..height - 2).contains(&cy);
            if inside {
                self.upload_info((width * cy + cx) as u64 * 16, &cell)?;
            }
        }
        Ok(())
    }

    /// `compute_by_pass(cpass, swap_index)` (d2q9_node.rs:302-312): collide_stream + boundary reading buffer
    /// `swap_index`, as one fused kernel.
    pub fn compute_by_pass(&self, swap_index: usize) -> LbmResult<()> {
        self.check(unsafe { sys::lbm_step(self.sim, swap_index as i32) })
    }

    /// `FluidSimulator::compute` (fluid_simulator.rs:217-232) inside the library: per frame ONE two-update sweep that
    /// also stores the macro texture of both updates, then the two particle passes (particles only read the field,
    /// so this is order-equivalent to step(0), particles, step(1), particles).
    pub fn compute_frames(&self, n_frames: i32) -> LbmResult<()> {
        self.check(unsafe { sys::lbm_compute_frames(self.sim, n_frames) })
    }

    /// `queue.write_buffer(&lbm_uniform_buf, ..)` (fluid_simulator.rs:188-192)
    pub fn write_uniform(&mut self, uniform: LbmUniform) -> LbmResult<()> {
        self.lbm_uniform_data = uniform;
        self.check(unsafe { sys::lbm_write_uniform(self.sim, (&self.lbm_uniform_data as *const LbmUniform).cast()) })
    }

    // ---- tracer particles and canvas (owned by the handle; fluid_simulator.rs:99-101 borrows them from the app)
    pub fn write_particle_uniform(&self, pu: &crate::ParticleUniform) -> LbmResult<()> {
        self.check(unsafe { sys::lbm_write_particle_uniform(self.sim, (pu as *const crate::ParticleUniform).cast()) })
    }

    pub fn write_particles(&self, particles: &[crate::TrajectoryParticle]) -> LbmResult<()> {
        self.check(unsafe { sys::lbm_particles_write(self.sim, particles.as_ptr().cast(), particles.len() as u64) })
    }

    /// The canvas `particle_update.wgsl` splats into, for the wgpu present pass (12 bytes per pixel).
    pub fn read_canvas(&self, dst: &mut [crate::Pixel]) -> LbmResult<()> {
        self.check(unsafe { sys::lbm_canvas_read(self.sim, dst.as_mut_ptr().cast()) })
    }

    /// present.wgsl:43-49 — the alpha fade the present pass applies to the canvas in place.
    pub fn fade_canvas(&self) -> LbmResult<()> {
        self.check(unsafe { sys::lbm_canvas_fade(self.sim) })
    }

    /// `macro_tex` contents (d2q9_node.rs:91-104): width * height RGBA16F texels.
    pub fn read_macro_tex(&self, dst: &mut [[u16; 4]]) -> LbmResult<()> {
        assert_eq!(dst.len(), (self.lattice.width * self.lattice.height) as usize);
        self.check(unsafe { sys::lbm_read_macro(self.sim, sys::LBM_MACRO_RGBA16F, dst.as_mut_ptr().cast()) })
    }

    /// `curl_tex` of the reference's `_curl_cal_node` (fluid_simulator.rs:36-71, curl_update.wgsl:12-33).
    pub fn read_curl_tex(&self, dst: &mut [[u16; 4]]) -> LbmResult<()> {
        assert_eq!(dst.len(), (self.lattice.width * self.lattice.height) as usize);
        self.check(unsafe { sys::lbm_read_curl(self.sim, dst.as_mut_ptr().cast()) })
    }

    /// Fragment outputs of the reference's `render_node` (fluid_simulator.rs:69-87, lbm/present.wgsl:21-46) for rows
    /// `row0 .. row0 + rows` of the canvas: `rows * canvas_width` RGBA f32 pixels.
    pub fn read_present(&self, row0: u32, rows: u32, canvas_width: u32, dst: &mut [[f32; 4]]) -> LbmResult<()> {
        assert_eq!(dst.len(), (rows * canvas_width) as usize);
        self.check(unsafe { sys::lbm_read_present(self.sim, row0 as i32, rows as i32, dst.as_mut_ptr().cast()) })
    }

    pub fn sync(&self) -> LbmResult<()> {
        self.check(unsafe { sys::lbm_sync(self.sim) })
    }
}

impl Drop for CudaD2Q9Node {
    fn drop(&mut self) {
        // SAFETY: `sim` came from lbm_create and is destroyed exactly once.
        unsafe { sys::lbm_destroy(self.sim) }
    }
}
