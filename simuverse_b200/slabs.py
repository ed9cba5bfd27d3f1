"""y-slab decomposition of one lattice over several GPUs.

The reference is single-GPU; this is the scale-out of its step.  Slab ``r`` of ``G`` owns rows
``[ny*r/G, ny*(r+1)/G)``.  There is no halo exchange call: after ``attach`` the first and last
row of every slab read / write the neighbour slab's memory directly over NVLink inside the
step kernel, ordered by device-side progress flags (csrc/lbm_step_vec.cuh).

Two ways to drive it:

* ``SlabRank`` — one process per GPU under ``torch.distributed`` (what bench.py uses);
  the 256-byte IPC blobs are exchanged with ``all_gather``.
* ``SlabGroup`` — all slabs in one process (several GPUs, or several slabs on one GPU for
  tests); steps are issued round-robin so no slab's kernel waits on an unlaunched neighbour.
"""
import numpy as np

from .d2q9_node import D2Q9Node


def slab_bounds(ny, rank, world):
    """Rows [y0, y1) owned by ``rank`` — same formula as lbm_create."""
    return ny * rank // world, ny * (rank + 1) // world


def neighbours(rank, world):
    """(up, down) ranks: owners of rows y0-1 and y1 with periodic wrap (layout_and_fn.wgsl:45-49)."""
    return (rank - 1) % world, (rank + 1) % world


def exchange_blobs(dist, mine, group=None, device=None):
    """all_gather of one fixed-size byte blob per rank; returns (all blobs, up blob, down blob).
    Works on any backend (gloo on CPU tensors, nccl on CUDA tensors)."""
    import torch

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        backend = dist.get_backend(group)
        device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(device)
    blobs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(blobs, t, group=group)
    out = [bytes(b.cpu().numpy()) for b in blobs]
    up, down = neighbours(rank, world)
    return out, out[up], out[down]


def gather_rows(dist, local, ny, group=None):
    """Assemble per-slab arrays of shape (..., rows_r, nx) into the global (..., ny, nx) on every rank
    (read-back path of a distributed run; off the hot path)."""
    import torch

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    rows = [slab_bounds(ny, r, world)[1] - slab_bounds(ny, r, world)[0] for r in range(world)]
    assert local.shape[-2] == rows[rank]
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    pad_shape = local.shape[:-2] + (max(rows), local.shape[-1])
    pad = torch.zeros(pad_shape, dtype=torch.from_numpy(local[..., :0, :]).dtype, device=dev)
    pad[..., :rows[rank], :] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return np.concatenate([parts[r][..., :rows[r], :].cpu().numpy() for r in range(world)], axis=-2)


class SlabRank:
    """This process' slab of a lattice decomposed over ``dist.get_world_size()`` ranks."""

    def __init__(self, canvas_size, setting, *, lattice, dist, device, group=None, **node_kwargs):
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.node = D2Q9Node(canvas_size, setting, lattice=lattice, device=device, rank=self.rank, world=self.world,
                             **node_kwargs)
        if self.world > 1:
            self._attach()
        self.barrier()
        self.node.reset()
        self.node.sync()
        self.barrier()

    def _attach(self):
        _, up, down = exchange_blobs(self.dist, self.node.ipc_export(), self.group)
        self.node.ipc_attach(up, down)

    def refresh_previous(self):
        """Collective: after two-update sweeps, bring the non-current buffer back to "one update behind" on every
        slab (lbm_refresh_previous) and wait until all slabs are done.  Call before reading that buffer or the
        on-demand macro field."""
        self.node.refresh_previous()
        self.barrier()

    def gather_distributions(self, which=None):
        """Global (9, ny, nx) distributions on every rank."""
        which = self.node.swap_index if which is None else which
        if which != self.node.swap_index:
            self.refresh_previous()
        return gather_rows(self.dist, self.node.read_distributions(which), self.node.lattice[1], self.group)

    def barrier(self):
        """Host-level rendezvous: every slab's queued work is done on return."""
        self.node.sync()
        if self.world > 1:
            self.dist.barrier(group=self.group)

    def step_n(self, n):
        self.node.step_n(n)

    def total_mass(self):
        """Global f64 mass: one all-reduce off the hot path."""
        import torch

        backend = self.dist.get_backend(self.group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        t = torch.tensor([self.node.total_mass()], dtype=torch.float64, device=dev)
        if self.world > 1:
            self.dist.all_reduce(t, group=self.group)
        return float(t.item())


class SlabGroup:
    """All slabs of one lattice driven from a single process."""

    def __init__(self, canvas_size, setting, *, lattice, n_slabs, devices=None, lattice_info=None, **node_kwargs):
        devices = list(devices) if devices is not None else [-1] * n_slabs
        assert len(devices) == n_slabs
        self.lattice = lattice
        self.nodes = [
            D2Q9Node(canvas_size, setting, lattice=lattice, lattice_info=lattice_info, device=devices[r], rank=r,
                     world=n_slabs, **node_kwargs)
            for r in range(n_slabs)
        ]
        if n_slabs > 1:
            blobs = [n.ipc_export() for n in self.nodes]
            for r, n in enumerate(self.nodes):
                up, down = neighbours(r, n_slabs)
                n.ipc_attach(blobs[up], blobs[down])
            self.reset()

    def sync(self):
        for n in self.nodes:
            n.sync()

    def reset(self):
        self.sync()
        for n in self.nodes:
            n.reset()
        self.sync()

    def step_n(self, n_steps):
        # round-robin: a slab's launch t waits (on the device) for its neighbours' launch t-1; pairs of updates go
        # out as one lbm_step_n(2) so that every slab may run them as one two-update sweep
        for _ in range(n_steps // 2):
            for n in self.nodes:
                n.step_n(2)
        if n_steps % 2:
            for n in self.nodes:
                n.step_n(1)

    def refresh_previous(self):
        """After sweeps: recompute the non-current buffer on every slab (collective), then wait for all of them."""
        for n in self.nodes:
            n.refresh_previous()
        self.sync()

    def write_lattice_info(self, byte_offset, cells):
        self.sync()
        self.refresh_previous()
        for n in self.nodes:
            n.write_lattice_info(byte_offset, cells)
        self.sync()

    @property
    def swap_index(self):
        return self.nodes[0].swap_index

    def read_distributions(self, which):
        self.sync()
        if which != self.swap_index:
            self.refresh_previous()
        return np.concatenate([n.read_distributions(which) for n in self.nodes], axis=1)

    def read_macro(self):
        self.sync()
        self.refresh_previous()  # the on-demand field is pulled from the buffer one update back
        return np.concatenate([n.read_macro() for n in self.nodes], axis=1)

    def read_lattice_info(self):
        self.sync()
        return np.concatenate([n.read_lattice_info() for n in self.nodes])

    def total_mass(self):
        self.sync()
        return sum(n.total_mass() for n in self.nodes)

    def close(self):
        self.sync()
        for n in self.nodes:
            n.close()
