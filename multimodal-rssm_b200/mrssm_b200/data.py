"""Host -> device input pipeline for the train step (the part the hot path needs).

Mirrors what ``ExperienceReplay_Multimodal.sample`` hands to ``optimize`` (utils/replay_buffer/memory.py:191-222):
frames live on the HOST as uint8 (memory.py:160-168); ``sample`` moves the gathered chunk to the device as uint8 and
converts / normalises there (memory.py:197-208, image_processing.py:5-11), returning time-major fp32
``[obs dict, actions, rewards, nonterminals]``.  Here the host buffers are pinned, the copy of batch k+1 runs on a side
stream while step k computes, and the normalisation is one kernel (``mrssm_normalize_image_u8``).  This source is handed
ready-made chunks (bench.py's e2e leg); the full replay buffer — chunk sampling, augmentation, device-resident uint8 frame
store — is ``utils/replay_buffer/memory.py``.
"""
import contextlib
import ctypes as C
import os

import torch

from . import _lib as L


def _device_local_cpus(device):
    """CPUs of the NUMA node the GPU hangs off (sysfs), or None when that cannot be determined."""
    try:
        p = torch.cuda.get_device_properties(device)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        cpus = set()
        for part in open(path).read().strip().split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        return cpus if cpus and cpus != allowed else None
    except Exception:
        return None


@contextlib.contextmanager
def near_device(device):
    """Run the body on the CPUs of the GPU's NUMA node: pinned buffers allocated (first touched) inside land in the memory
    the GPU's PCIe root complex reaches without crossing the socket interconnect.  No-op when the topology is unknown."""
    cpus = _device_local_cpus(device)
    if cpus is None:
        yield
        return
    old = os.sched_getaffinity(0)
    try:
        os.sched_setaffinity(0, cpus)
        yield
    finally:
        os.sched_setaffinity(0, old)


class PinnedChunkSource:
    """D.sample(n, L) over a ring of pre-gathered host chunks.

    chunks: list of (obs dict name -> tensor [T,B,...] (uint8 for images, fp32 otherwise), actions [T,B,A],
    rewards [T,B], nonterminals [T,B,1]) on the CPU; they are pinned here."""

    def __init__(self, chunks, device, bit_depth=5, seed=0, prefetch=True):
        self.device = torch.device(device)
        self.bit_depth, self.seed, self.prefetch = bit_depth, seed, prefetch
        pin = lambda t: t.contiguous().pin_memory()
        with near_device(self.device):       # pinned staging memory NUMA-local to the GPU
            self.host = [({k: pin(v) for k, v in obs.items()}, pin(a), pin(r), pin(n)) for obs, a, r, n in chunks]
        obs, a, r, n = self.host[0]
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in obs.values()) + 4 * (a.numel() + r.numel() + n.numel())
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = [None, None]          # device staging, double buffered
        self._s2d_ready = [False, False]
        self.i = 0
        self.pending = None                # (slot index, ready event) of the batch in flight

    def _slot(self, k):
        if self.slots[k] is None:
            obs, a, r, n = self.host[0]
            dev = self.device
            raw = {name: torch.empty_like(v, device=dev) for name, v in obs.items()}
            f32 = {name: (torch.empty(v.shape, device=dev, dtype=torch.float32) if v.dtype == torch.uint8 else raw[name])
                   for name, v in obs.items()}
            # bf16 mode: the space-to-depth bf16 form of frames 1.. of every <= 4-channel image (what the conv encoder's first layer
            # and the fused reconstruction loss read; optimize() drops frame 0, base/algo.py:213-215), made on the copy stream too
            s2d = {}
            for name, v in obs.items():
                if v.dtype == torch.uint8 and v.dim() == 5 and v.shape[2] <= 4 and v.shape[0] > 1 and v.shape[3] % 2 == 0 and v.shape[4] % 2 == 0:
                    T, B, Cc, H, W = v.shape
                    s2d[name] = torch.empty(L.view_numel(L.PLANAR, (T - 1) * B, H // 2, W // 2, 16), device=dev, dtype=torch.bfloat16)
            self.slots[k] = (raw, f32, torch.empty_like(a, device=dev), torch.empty_like(r, device=dev),
                             torch.empty_like(n, device=dev), s2d)
        return self.slots[k]

    def _launch(self, k, chunk_index):
        """Enqueue H2D + normalisation of one chunk into slot k on the copy stream."""
        self._s2d_ready[k] = False
        obs, a, r, n = self.host[chunk_index % len(self.host)]
        raw, f32, da, dr, dn, s2d = self._slot(k)
        from . import ops
        main = torch.cuda.current_stream(self.device)
        self.copy_stream.wait_stream(main)                 # the slot's previous consumer has been enqueued before us
        with torch.cuda.stream(self.copy_stream):
            for name, v in obs.items():
                raw[name].copy_(v, non_blocking=True)
                if v.dtype == torch.uint8:
                    L.call("mrssm_normalize_image_u8", L.ptr_any(raw[name]), v.numel(), self.bit_depth, None,
                           self.seed + chunk_index, L.ptr(f32[name]))
                    if name in s2d and ops.bf16_mode():
                        T, B, Cc, H, W = v.shape
                        x1 = f32[name][1:]
                        L.call("mrssm_pl_import_s2d", C.byref(L.nchw(x1, H, W, Cc)), (T - 1) * B, H, W, Cc, 1.0,
                               C.byref(L.tv(s2d[name], L.PLANAR, H // 2, W // 2, 16)))
                        self._s2d_ready[k] = True
            da.copy_(a, non_blocking=True)
            dr.copy_(r, non_blocking=True)
            dn.copy_(n, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return ev

    def sample(self, n, L_):
        k = self.i & 1
        if self.pending is None or self.pending[0] != k:
            self.pending = (k, self._launch(k, self.i))
        _, ev = self.pending
        torch.cuda.current_stream(self.device).wait_event(ev)
        raw, f32, da, dr, dn, s2d = self.slots[k]
        if self._s2d_ready[k]:
            from . import ops
            for name, t in s2d.items():    # keyed like the tensor the encoder will see: frames 1.. folded to [(T-1)*B, C, H, W]
                x1 = f32[name][1:]
                ops.remember_s2d(x1.reshape(-1, *x1.shape[2:]), t)
        self.i += 1
        self.pending = (k ^ 1, self._launch(k ^ 1, self.i)) if self.prefetch else None
        return [dict(f32), da, dr, dn]
