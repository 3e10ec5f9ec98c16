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


class HostStepPipeline:
    """Streaming form of step_from_host for a caller that feeds one batch of particle
    attributes per step from pinned host memory and reads every step's result back:

        pipe = HostStepPipeline(loop, electrons)
        for k in range(K):
            pipe.submit(host_in, host_out[k % 2])      # never blocks the host
        pipe.drain()

    Every step still uploads all its inputs and downloads all its results; what changes
    is WHEN: the device keeps two sets of attribute arrays, so the upload of step k+1
    (copy engine 1, 2.1 GB at cfg3) runs while step k computes and while the results of
    step k go back (copy engine 2, 1.9 GB) -- both PCIe directions are busy at the same
    time and the step costs max(upload, download) instead of their sum.  Inside a step
    the coordinates leave as soon as the one-pass particle side is done (they are final
    then), the momenta after the gather.  The caller must not touch host_out of step k
    before step k+2 has been submitted and `wait(k)` returned (or after drain())."""

    def __init__(self, loop, parts, depth=2):
        self.loop, self.parts = loop, parts
        dev = parts.comm.device
        self.main = torch.cuda.current_stream()
        self.up = torch.cuda.Stream(device=dev)
        self.down = torch.cuda.Stream(device=dev)
        first = {a: parts.DataDev[a] for a in ATTRS_IN}
        from .devarray import DevArray
        self.sets = [first] + [{a: DevArray(torch.empty_like(first[a].t)) for a in ATTRS_IN}
                               for _ in range(depth - 1)]
        self.free = [None] * depth          # event: the set may be overwritten again
        self.done = {}                      # step -> event: host_out of that step is valid
        self.rho_read = None
        self.k = 0

    def submit(self, host_in, host_out):
        k, parts, loop = self.k, self.parts, self.loop
        cur = self.sets[k % len(self.sets)]
        h2d = d2h = 0
        # upload on its own stream, after the previous user of this set has been read back
        if self.free[k % len(self.sets)] is not None:
            self.up.wait_event(self.free[k % len(self.sets)])
        else:
            self.up.wait_stream(self.main)
        with torch.cuda.stream(self.up):
            for a in ATTRS_IN:
                cur[a].t.copy_(host_in[a], non_blocking=True)
                h2d += host_in[a].numel() * 8
            uploaded = torch.cuda.Event()
            uploaded.record(self.up)
        self.main.wait_event(uploaded)
        if self.rho_read is not None:
            self.main.wait_event(self.rho_read)
        for a in ATTRS_IN:
            parts.DataDev[a] = cur[a]
        parts.flag_sorted = False

        def copy_coordinates(_loop):
            ready = torch.cuda.Event()
            ready.record(self.main)
            self.down.wait_event(ready)
            with torch.cuda.stream(self.down):
                for a in ("x", "y", "z"):
                    host_out[a].copy_(cur[a].t, non_blocking=True)

        loop.on_coordinates_final = copy_coordinates
        try:
            loop.step()
        finally:
            loop.on_coordinates_final = None
        computed = torch.cuda.Event()
        computed.record(self.main)
        self.down.wait_event(computed)
        rho = loop.mainsolver.DataDev["rho_m0"].t
        with torch.cuda.stream(self.down):
            # rho first: the next step's charge deposit rewrites it, and that step may
            # start as soon as this one has computed (its upload is already there)
            host_out["rho_m0"].copy_(rho, non_blocking=True)
            d2h += host_out["rho_m0"].numel() * 8
            self.rho_read = torch.cuda.Event()
            self.rho_read.record(self.down)
            for a in ATTRS_OUT:
                if a not in ("x", "y", "z"):
                    host_out[a].copy_(cur[a].t, non_blocking=True)
                d2h += host_out[a].numel() * 8
            back = torch.cuda.Event()
            back.record(self.down)
        self.free[k % len(self.sets)] = back
        self.done[k] = back
        self.done.pop(k - 4, None)
        self.k += 1
        return h2d, d2h

    def wait(self, k):
        self.done[k].synchronize()

    def drain(self):
        """Make the caller's stream wait for everything in flight (no host block)."""
        self.main.wait_stream(self.up)
        self.main.wait_stream(self.down)


_SIDE = {}


def _side_stream(device):
    key = str(device)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]
