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


def step_from_host(loop, parts, host_in, host_out, overlap=True):
    """One PIC step with the mobile species' attributes coming from / going back to
    pinned host memory.  overlap=True: the coordinates are final after the one-pass
    particle side (about 1 ms into the step), so their download runs on a second stream
    underneath the field solve and the gather; the momenta follow on the main stream."""
    solver = loop.mainsolver
    main = torch.cuda.current_stream()
    h2d = 0
    for a in ATTRS_IN:
        parts.DataDev[a].t.copy_(host_in[a], non_blocking=True)
        h2d += host_in[a].numel() * 8
    parts.flag_sorted = False
    early = ("x", "y", "z") if overlap else ()
    done = []

    def copy_coordinates(_loop):
        side = _side_stream(parts.comm.device)
        ready = torch.cuda.Event()
        ready.record(main)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            for a in early:
                host_out[a].copy_(parts.DataDev[a].t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(side)
        done.append(ev)

    loop.on_coordinates_final = copy_coordinates if overlap else None
    try:
        loop.step()
    finally:
        loop.on_coordinates_final = None
    d2h = 0
    for a in ATTRS_OUT:
        if a not in early:
            host_out[a].copy_(parts.DataDev[a].t, non_blocking=True)
        d2h += host_out[a].numel() * 8
    host_out["rho_m0"].copy_(solver.DataDev["rho_m0"].t, non_blocking=True)
    d2h += host_out["rho_m0"].numel() * 8
    for ev in done:                      # the step is complete when both streams are
        main.wait_event(ev)
    return h2d, d2h


_SIDE = {}


def _side_stream(device):
    key = str(device)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]
