// Compiles the C++ host mirror (include/d2q9_node.hpp) against liblbm_b200.so.
// Without a GPU: checks the host helpers and that construction fails loudly (no CPU fallback).
// With a GPU: runs two frames of FluidSimulator::compute, a click and a drag, prints a checksum.
#include <cstdio>
#include <cstring>
#include <vector>

#include "d2q9_node.hpp"

int main() {
    LbmUniform u;
    lbm_uniform_new(0.56f, 0, 225000, &u);
    if (u.omega != 1.0f / 0.56f || u.e_w_max[5][2] != 0.0277777f) return 2;
    std::vector<LatticeInfo> info(600 * 375);
    lbm_init_lattice_material(600, 375, FIELD_ANIMATION_POISEUILLE, info.data());
    long obstacles = 0;
    for (const auto &c : info) obstacles += c.material == LATTICE_OBSTACLE;
    if (obstacles != 7374) return 3;
    if (lbm_device_count() == 0) {
        try {
            lbm::D2Q9Node node({1200, 750}, lbm::SettingObj{});
            return 4; // must not succeed on the CPU
        } catch (const lbm::Error &e) {
            if (e.status != LBM_ERR_NO_DEVICE) return 5;
            std::printf("HOST_MIRROR_OK no-gpu (%s)\n", e.what());
            return 0;
        }
    }
    lbm::FluidSimulator sim({1200, 750}, lbm::SettingObj{});
    sim.compute();
    if (!sim.on_click({700.0f, 400.0f})) return 6;
    sim.touch_begin();
    sim.touch_move({300.0f, 300.0f});
    sim.touch_move({340.0f, 310.0f});
    sim.compute();
    double mass = 0.0;
    lbm::check(lbm_total_mass(sim.fluid_compute_node().handle(), lbm_swap_index(sim.fluid_compute_node().handle()), &mass));
    // the same session with every frame issued call by call (single-update kernels): identical distributions
    lbm::FluidSimulator ref({1200, 750}, lbm::SettingObj{});
    ref.compute_by_passes();
    ref.on_click({700.0f, 400.0f});
    ref.touch_begin();
    ref.touch_move({300.0f, 300.0f});
    ref.touch_move({340.0f, 310.0f});
    ref.compute_by_passes();
    std::vector<float> a(9 * 600 * 375), b(a.size());
    LbmSim *ha = sim.fluid_compute_node().handle(), *hb = ref.fluid_compute_node().handle();
    lbm::check(lbm_read_distributions(ha, lbm_swap_index(ha), a.data()), ha);
    lbm::check(lbm_read_distributions(hb, lbm_swap_index(hb), b.data()), hb);
    if (std::memcmp(a.data(), b.data(), a.size() * sizeof(float)) != 0) return 8;
    if (lbm_fused_sweep_count(ha) == 0 || lbm_fused_sweep_count(hb) != 0) return 9;
    std::printf("HOST_MIRROR_OK gpu mass=%.6f launches=%llu sweeps=%llu\n", mass,
                (unsigned long long)lbm_launch_count(ha), (unsigned long long)lbm_fused_sweep_count(ha));
    return mass > 2.0e5 && mass < 2.2e5 ? 0 : 7;
}
