"""End-to-end entry with HOST buffers: what a caller that keeps its particles in host
memory (the reference on a CPU OpenCL device) would invoke per step.

    step_from_host(loop, electrons, host_in, host_out)

uploads the mobile species' attribute arrays from pinned host memory, runs one full
PIC_loop.step() on the device and downloads the updated attributes plus rho_m0."""
import torch

ATTRS_IN = ("x", "y", "z", "px", "py", "pz", "w", "g_inv")
ATTRS_OUT = ("x", "y", "z", "px", "py", "pz", "g_inv")


def make_host_buffers(parts, solver):
    host_in = {a: parts.DataDev[a].t.cpu().pin_memory() for a in ATTRS_IN}
    host_out = {a: torch.empty_like(host_in[a]).pin_memory() for a in ATTRS_OUT}
    host_out["rho_m0"] = torch.empty(solver.DataDev["rho_m0"].shape, dtype=torch.float64).pin_memory()
    return host_in, host_out


def step_from_host(loop, parts, host_in, host_out):
    solver = loop.mainsolver
    h2d = 0
    for a in ATTRS_IN:
        parts.DataDev[a].t.copy_(host_in[a], non_blocking=True)
        h2d += host_in[a].numel() * 8
    parts.flag_sorted = False
    loop.step()
    d2h = 0
    for a in ATTRS_OUT:
        host_out[a].copy_(parts.DataDev[a].t, non_blocking=True)
        d2h += host_out[a].numel() * 8
    host_out["rho_m0"].copy_(solver.DataDev["rho_m0"].t, non_blocking=True)
    d2h += host_out["rho_m0"].numel() * 8
    return h2d, d2h
