//! `CudaFluidSimulator` — the LBM player's `impl Simulator` (simuverse/src/lib.rs:72-106) with the compute half of
//! `FluidSimulator` (simuverse/src/fluid/fluid_simulator.rs:14-249) moved onto the B200-native library.
//!
//! File placement: `simuverse/src/fluid/cuda_fluid_simulator.rs`; `fluid/mod.rs` gains
//! `mod cuda_d2q9_node; mod cuda_fluid_simulator; pub use cuda_fluid_simulator::CudaFluidSimulator;` and
//! `SimuverseApp::create_simulator` (simuverse_app.rs) constructs this type where it builds `FluidSimulator` today.
//! Rendering stays wgpu: the only render node the reference draws for this player is `particle_render`
//! (fluid_simulator.rs:234-248, present.wgsl over `canvas_buf`), so each frame the canvas the CUDA particle pass
//! splatted into is copied into the app's `canvas_buf`.  (A CUDA <-> Vulkan external-memory import of that buffer
//! would remove the copy; it is outside the compute path this replacement covers.)
//!
//! UNBUILT in this repository (no Rust toolchain in the image).  The same call sequence is exercised through the C ABI
//! by tests/test_gpu_parity.py::test_frame_loop_with_particles_matches_oracle (Python mirror) and
//! tests/cpp/host_mirror_check.cpp (C++ mirror).
use alloc::{vec, vec::Vec};

use super::{OBSTACLE_RADIUS, cuda_d2q9_node::CudaD2Q9Node};
use crate::{
    FieldAnimationType, Pixel, SettingObj, Simulator,
    fluid::LbmUniform,
    node::{BindGroupData, BufferlessFullscreenNode},
    util::BufferObj,
};

pub struct CudaFluidSimulator {
    lattice: wgpu::Extent3d,
    lattice_pixel_size: u32,
    pre_pos: glam::Vec2,
    fluid_compute_node: CudaD2Q9Node,
    particle_render: BufferlessFullscreenNode,
    // the app's canvas buffer (what present.wgsl reads) and the host staging copy of the CUDA canvas
    queue: wgpu::Queue,
    canvas_buf: wgpu::Buffer,
    canvas_host: Vec<Pixel>,
}

impl CudaFluidSimulator {
    /// `FluidSimulator::new` (fluid_simulator.rs:26-133) without the compute pipelines.
    pub fn new(
        app: &app_surface::AppSurface,
        canvas_size: glam::UVec2,
        canvas_buf: &BufferObj,
        setting: &SettingObj,
        particles: &[crate::TrajectoryParticle],
    ) -> Self {
        let grid = setting.particles_uniform_data.num;
        let mut fluid_compute_node = CudaD2Q9Node::new(app, canvas_size, setting, grid[0] * grid[1])
            .unwrap_or_else(|e| panic!("{e}")); // the reference panics on device errors too (util/shader.rs:81)
        fluid_compute_node
            .write_particle_uniform(&setting.particles_uniform_data)
            .and_then(|_| fluid_compute_node.write_particles(particles))
            .unwrap_or_else(|e| panic!("{e}"));

        // fluid_simulator.rs:114-127: the canvas present pass, unchanged
        let field_uniform_buf =
            BufferObj::create_uniform_buffer(&app.device, &fluid_compute_node.field_uniform_data, Some("fluid uniform"));
        let particle_shader = crate::create_shader_module(&app.device, "present", None);
        let particle_render = BufferlessFullscreenNode::new(
            &app.device,
            app.config.format,
            &BindGroupData {
                uniforms: vec![&field_uniform_buf, setting.particles_uniform.as_ref().unwrap()],
                storage_buffers: vec![canvas_buf],
                ..Default::default()
            },
            &particle_shader,
            None,
        );

        CudaFluidSimulator {
            lattice: fluid_compute_node.lattice,
            lattice_pixel_size: fluid_compute_node.lattice_pixel_size,
            pre_pos: glam::Vec2::ZERO,
            fluid_compute_node,
            particle_render,
            queue: app.queue.clone(),
            canvas_buf: canvas_buf.buffer.clone(),
            canvas_host: vec![Pixel { alpha: 0.0, speed: 0.0, rho: 0.0 }; (canvas_size.x * canvas_size.y) as usize],
        }
    }
}

impl Simulator for CudaFluidSimulator {
    /// fluid_simulator.rs:137-152
    fn on_click(&mut self, _app: &app_surface::AppSurface, pos: glam::Vec2) {
        if pos.x <= 0.0 || pos.y <= 0.0 {
            return;
        }
        let (x, y) = (pos.x as u32 / self.lattice_pixel_size, pos.y as u32 / self.lattice_pixel_size);
        let margin = OBSTACLE_RADIUS as u32;
        let x_ok = (margin..self.lattice.width - (margin + 2)).contains(&x);
        let y_ok = (margin..self.lattice.height - (margin + 2)).contains(&y);
        if x_ok && y_ok {
            self.fluid_compute_node.add_obstacle(x, y).unwrap_or_else(|e| panic!("{e}"));
        }
    }

    /// fluid_simulator.rs:154-156
    fn touch_begin(&mut self, _app: &app_surface::AppSurface) {
        self.pre_pos = glam::Vec2::ZERO;
    }

    /// fluid_simulator.rs:158-173
    fn touch_move(&mut self, _app: &app_surface::AppSurface, pos: glam::Vec2) {
        if pos.x <= 0.0 || pos.y <= 0.0 {
            self.pre_pos = glam::Vec2::ZERO;
            return;
        }
        let fresh_stroke = self.pre_pos == glam::Vec2::ZERO || pos.distance(self.pre_pos) > 300.0;
        if !fresh_stroke {
            self.fluid_compute_node
                .add_external_force(pos, self.pre_pos)
                .unwrap_or_else(|e| panic!("{e}"));
        }
        self.pre_pos = pos;
    }

    /// fluid_simulator.rs:175-193: tau = 3 * viscosity + 0.5, uniform uploaded again
    fn update_uniforms(&mut self, _app: &app_surface::AppSurface, setting: &crate::SettingObj) {
        let fluid_ty = (setting.animation_type == FieldAnimationType::LidDrivenCavity) as i32;
        let uniform = LbmUniform::new(
            3.0 * setting.fluid_viscosity + 0.5,
            fluid_ty,
            (self.lattice.width * self.lattice.height) as i32,
        );
        self.fluid_compute_node.write_uniform(uniform).unwrap_or_else(|e| panic!("{e}"));
    }

    fn update_by(&mut self, _app: &app_surface::AppSurface, _control_panel: &mut crate::ControlPanel) {}

    /// fluid_simulator.rs:203-208 sets the dispatch size of the particle pass; the library derives its grid from
    /// `ParticleUniform::num`, so there is nothing to store.
    fn update_workgroup_count(&mut self, _app: &app_surface::AppSurface, _workgroup_count: (u32, u32, u32)) {}

    /// fluid_simulator.rs:210-215
    fn reset(&mut self, _app: &app_surface::AppSurface) {
        self.fluid_compute_node.reset_lattice_info().unwrap_or_else(|e| panic!("{e}"));
        self.pre_pos = glam::Vec2::ZERO;
    }

    /// fluid_simulator.rs:217-232: one frame = update, particles, update, particles.  Nothing is recorded into the
    /// wgpu encoder: the frame runs on the library's CUDA stream, then the canvas goes to the present pass' buffer.
    fn compute(&mut self, _encoder: &mut wgpu::CommandEncoder) {
        let node = &self.fluid_compute_node;
        node.compute_frames(1)
            .and_then(|_| node.read_canvas(&mut self.canvas_host))
            .unwrap_or_else(|e| panic!("{e}"));
        self.queue.write_buffer(&self.canvas_buf, 0, bytemuck::cast_slice(&self.canvas_host));
    }

    /// fluid_simulator.rs:234-248: present.wgsl draws the canvas and fades it in place (present.wgsl:43-49); the fade
    /// is applied to the CUDA-side canvas as well, which is the copy the next frame's particle passes continue from.
    fn draw_by_rpass<'b, 'a: 'b>(
        &'a mut self,
        _app: &app_surface::AppSurface,
        rpass: &mut wgpu::RenderPass<'b>,
        _setting: &mut crate::SettingObj,
    ) {
        self.particle_render.draw_by_pass(rpass);
        self.fluid_compute_node.fade_canvas().unwrap_or_else(|e| panic!("{e}"));
    }
}
