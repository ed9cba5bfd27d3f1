// d2q9_node.hpp — C++ host-side mirror of the reference's LBM player over the C ABI
// (include/lbm_b200.h).  Header-only; link against liblbm_b200.so.
//
//   lbm::D2Q9Node        simuverse/src/fluid/d2q9_node.rs:30-313
//   lbm::FluidSimulator  simuverse/src/fluid/fluid_simulator.rs:14-249  (impl Simulator, lib.rs:72-106)
//
// Same entry points and argument meaning as the Rust types; the wgpu Device/Queue/CommandEncoder
// parameters have no counterpart (calls are stream-ordered on the handle), and errors surface as
// lbm::Error instead of panics.
#pragma once

#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "lbm_b200.h"

namespace lbm {

struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string &what) : std::runtime_error(what), status(st) {}
};

inline void check(int st, const LbmSim *sim = nullptr) {
    if (st != LBM_OK) throw Error(st, std::string(lbm_status_string(st)) + ": " + lbm_last_error(sim));
}

struct Vec2 {
    float x = 0.0f, y = 0.0f;
};

// The slice of SettingObj (setting/setting_obj.rs:4-56) the LBM player reads.
struct SettingObj {
    int32_t animation_type = FIELD_ANIMATION_POISEUILLE;
    float fluid_viscosity = 0.02f;  // setting_obj.rs:32
    int32_t particles_count = 10000; // control_panel.rs:29
    ParticleUniform particles_uniform_data{{1.f, 1.f, 1.f, 1.f}, {0, 0}, 2, 90.0f, 0.96f, 4.15f, 0, 1};
};

class D2Q9Node {
  public:
    std::pair<int32_t, int32_t> lattice;        // (nx, ny)
    uint32_t lattice_pixel_size;
    std::vector<LatticeInfo> lattice_info_data; // CPU mirror, NOT updated by force writes (like the reference)
    LbmUniform lbm_uniform_data{};
    FieldUniform field_uniform_data{};

    // D2Q9Node::new (d2q9_node.rs:31-209). `device` = CUDA ordinal (-1: current).
    D2Q9Node(std::pair<uint32_t, uint32_t> canvas_size, const SettingObj &setting, float scale_factor = 1.0f,
             uint32_t flags = 0, int32_t max_particles = 0, int32_t device = -1)
        : lattice_pixel_size(static_cast<uint32_t>(std::ceil(2.0f * scale_factor))), animation_ty_(setting.animation_type) {
        lattice = {static_cast<int32_t>(canvas_size.first / lattice_pixel_size),
                   static_cast<int32_t>(canvas_size.second / lattice_pixel_size)};
        LbmDesc d{};
        d.struct_size = sizeof(LbmDesc);
        d.nx = lattice.first;
        d.ny = lattice.second;
        d.lattice_pixel_size = static_cast<int32_t>(lattice_pixel_size);
        d.canvas_w = static_cast<int32_t>(canvas_size.first);
        d.canvas_h = static_cast<int32_t>(canvas_size.second);
        d.device = device;
        d.rank = 0;
        d.world = 1;
        d.flags = flags;
        d.max_particles = max_particles;
        check(lbm_create(&d, &sim_));
        const float tau = lbm_tau_from_viscosity(setting.fluid_viscosity); // d2q9_node.rs:50
        const int32_t fluid_ty = setting.animation_type == FIELD_ANIMATION_LID_DRIVEN_CAVITY ? 1 : 0;
        lbm_uniform_new(tau, fluid_ty, lattice.first * lattice.second, &lbm_uniform_data);
        check(lbm_write_uniform(sim_, &lbm_uniform_data), sim_);
        lbm_field_uniform_new(lattice.first, lattice.second, lattice_pixel_size, d.canvas_w, d.canvas_h, &field_uniform_data);
        check(lbm_write_field_uniform(sim_, &field_uniform_data), sim_);
        lattice_info_data.resize(static_cast<size_t>(lattice.first) * lattice.second);
        lbm_init_lattice_material(lattice.first, lattice.second, animation_ty_, lattice_info_data.data()); // :106
        write_info(0, lattice_info_data.data(), lattice_info_data.size());
        reset(); // :206
    }
    ~D2Q9Node() { lbm_destroy(sim_); }
    D2Q9Node(const D2Q9Node &) = delete;
    D2Q9Node &operator=(const D2Q9Node &) = delete;

    LbmSim *handle() { return sim_; }

    // d2q9_node.rs:211-213
    void reset() { check(lbm_reset(sim_), sim_); }

    // d2q9_node.rs:215-245
    void add_obstacle(uint32_t x, uint32_t y) {
        std::vector<LatticeInfo> patch(static_cast<size_t>(56) * lattice.first);
        uint64_t off = 0;
        const uint64_t n = lbm_obstacle_patch(lattice.first, lattice.second, lattice_info_data.data(), x, y, patch.data(), &off);
        check(lbm_write_lattice_info(sim_, off, patch.data(), n * sizeof(LatticeInfo)), sim_);
    }

    // d2q9_node.rs:247-261
    void reset_lattice_info() {
        if (animation_ty_ == FIELD_ANIMATION_POISEUILLE) {
            lbm_init_lattice_material(lattice.first, lattice.second, animation_ty_, lattice_info_data.data());
            write_info(0, lattice_info_data.data(), lattice_info_data.size());
        }
        reset();
    }

    // d2q9_node.rs:263-300
    void add_external_force(Vec2 pos, Vec2 pre_pos) {
        uint64_t offs[4096];
        std::vector<LatticeInfo> cells(4096);
        uint64_t n = lbm_external_force_cells(lattice.first, lattice.second, lattice_pixel_size, pos.x, pos.y, pre_pos.x,
                                              pre_pos.y, offs, cells.data(), 4096);
        if (n > 4096) n = 4096;
        for (uint64_t k = 0; k < n; k++) check(lbm_write_lattice_info(sim_, offs[k], &cells[k], sizeof(LatticeInfo)), sim_);
    }

    // d2q9_node.rs:302-312 (collide_stream + boundary of bind group `swap_index`, one fused kernel here)
    void compute_by_pass(int32_t swap_index) { check(lbm_step(sim_, swap_index), sim_); }

    void write_uniform(const LbmUniform &u) {
        lbm_uniform_data = u;
        check(lbm_write_uniform(sim_, &lbm_uniform_data), sim_);
    }

  private:
    void write_info(uint64_t first_cell, const LatticeInfo *cells, size_t count) {
        check(lbm_write_lattice_info(sim_, first_cell * sizeof(LatticeInfo), cells, count * sizeof(LatticeInfo)), sim_);
    }
    LbmSim *sim_ = nullptr;
    int32_t animation_ty_;
};

class FluidSimulator {
  public:
    static constexpr uint32_t kObstacleRadius = 28; // fluid/mod.rs:1

    // FluidSimulator::new (fluid_simulator.rs:26-133); `particle_seed` replaces the unseeded rand::rng().
    FluidSimulator(std::pair<uint32_t, uint32_t> canvas_size, SettingObj setting, float scale_factor = 1.0f,
                   uint64_t particle_seed = 0x5EED, int32_t device = -1)
        : setting_(setting), num_(grid(canvas_size, setting.particles_count)),
          node_(canvas_size, setting, scale_factor, LBM_FLAG_MACRO_EVERY_STEP, num_.first * num_.second, device) {
        setting_.particles_uniform_data.num[0] = num_.first;
        setting_.particles_uniform_data.num[1] = num_.second;
        check(lbm_write_particle_uniform(node_.handle(), &setting_.particles_uniform_data), node_.handle());
        std::vector<TrajectoryParticle> p(static_cast<size_t>(num_.first) * num_.second);
        lbm_init_trajectory_particles(canvas_size.first, canvas_size.second, num_.first, num_.second,
                                      setting_.particles_uniform_data.life_time, particle_seed, p.data());
        check(lbm_particles_write(node_.handle(), p.data(), p.size()), node_.handle());
    }

    D2Q9Node &fluid_compute_node() { return node_; }

    // fluid_simulator.rs:137-152
    bool on_click(Vec2 pos) {
        uint32_t x = 0, y = 0;
        if (!lbm_on_click_guard(node_.lattice.first, node_.lattice.second, node_.lattice_pixel_size, pos.x, pos.y, &x, &y))
            return false;
        node_.add_obstacle(x, y);
        return true;
    }
    // fluid_simulator.rs:154-156
    void touch_begin() { pre_pos_ = {}; }
    // fluid_simulator.rs:158-173
    void touch_move(Vec2 pos) {
        if (pos.x <= 0.0f || pos.y <= 0.0f) {
            pre_pos_ = {};
            return;
        }
        const float dx = pos.x - pre_pos_.x, dy = pos.y - pre_pos_.y;
        const float dis = std::sqrt(dx * dx + dy * dy);
        if ((pre_pos_.x == 0.0f && pre_pos_.y == 0.0f) || dis > 300.0f) {
            pre_pos_ = pos;
            return;
        }
        node_.add_external_force(pos, pre_pos_);
        pre_pos_ = pos;
    }
    // fluid_simulator.rs:175-193
    void update_uniforms(const SettingObj &setting) {
        LbmUniform u;
        lbm_uniform_new(lbm_tau_from_viscosity(setting.fluid_viscosity),
                        setting.animation_type == FIELD_ANIMATION_LID_DRIVEN_CAVITY ? 1 : 0,
                        node_.lattice.first * node_.lattice.second, &u);
        node_.write_uniform(u);
    }
    // fluid_simulator.rs:210-215
    void reset() {
        node_.reset_lattice_info();
        pre_pos_ = {};
    }
    // fluid_simulator.rs:217-232: one frame = step(0), particles, step(1), particles — inside the library ONE
    // two-update sweep that stores the macro texture of both updates, then the two particle passes (same results)
    void compute(int32_t n_frames = 1) { check(lbm_compute_frames(node_.handle(), n_frames), node_.handle()); }
    // the same frame issued call by call, as the reference records it (single-update kernels)
    void compute_by_passes() {
        node_.compute_by_pass(0);
        check(lbm_particles_update(node_.handle()), node_.handle());
        node_.compute_by_pass(1);
        check(lbm_particles_update(node_.handle()), node_.handle());
    }

    // fluid_simulator.rs:234-248: only the state-changing part of the present pass (canvas alpha fade)
    void draw_by_rpass() { check(lbm_canvas_fade(node_.handle()), node_.handle()); }

  private:
    static std::pair<int32_t, int32_t> grid(std::pair<uint32_t, uint32_t> canvas, int32_t count) {
        int32_t a = 0, b = 0;
        lbm_particle_grid(canvas.first, canvas.second, count, &a, &b);
        return {a, b};
    }
    SettingObj setting_;
    std::pair<int32_t, int32_t> num_;
    D2Q9Node node_;
    Vec2 pre_pos_{};
};

}  // namespace lbm
