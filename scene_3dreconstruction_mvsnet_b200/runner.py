"""Depth-map runner: the user-facing call for inference with HOST buffers.

Mirrors what the reference's eval loop does around the model call (eval.py:358-360: tocuda(sample)
-> model(...) -> tensor2numpy(outputs)), but with pinned staging buffers, a dedicated copy stream and
double buffering so the host->device copy of view i+1 and the device->host copy of view i-1 overlap
the kernels of view i.  Every byte still crosses PCIe inside the call -- this is the path bench.py
times as `e2e`.
"""
import itertools

import numpy as np
import torch


class DepthMapRunner:
    """graphs=True (tensor-core modes): the forward pass of every staging slot is captured once into a CUDA graph --
    the slot's device input buffers and its output tensor are static -- and replayed for later views.  One graph launch
    replaces ~25 kernel launches with their Python / ctypes / tensor-map set-up (0.43 ms of host time per depth map).
    Measured on an idle host (profiles/r02t): no gain -- the eager path already keeps the GPU busy (512 x 640, 4 views:
    1834 eager vs 1814 depth maps/s replayed; 1152 x 1600: 416 vs 399) -- so it is off by default; it is for hosts whose
    cores are contended (8 ranks per node) or slow."""

    def __init__(self, model, device="cuda:0", depth=3, graphs=False, streams="auto"):
        """streams: compute streams the views alternate over (each with its own workspaces).  Consecutive depth maps are
        independent, so the next map's kernels fill the SMs that the tail of a kernel leaves idle: +3 % end to end at
        1152 x 1600 (e2e 418 -> 430, uint8 432 -> 447 depth maps/s).  At 512 x 640 the host's launch work is the bound
        and a second stream costs 3-9 %, so "auto" uses two streams from 2 M input pixels per depth map on;
        1 = everything on the caller's current stream."""
        self.model = model.to(device).eval()
        self.device = torch.device(device)
        self.n_streams = streams if streams == "auto" else max(1, int(streams))
        self._compute_streams = None
        if graphs:
            self.n_streams = 1   # the captured graphs share one set of workspaces: replays must not overlap
        self.graphs = bool(graphs) and getattr(model, "precision", "fp32") in ("bf16", "fast")
        self._capture_stream = None
        self._graph_pool = None
        # separate streams (and copy engines) for uploads and downloads: a download waits for its forward pass, and
        # on a shared stream it would hold back the next view's upload until that forward pass is over
        self.copy_stream = torch.cuda.Stream(self.device)
        self.d2h_stream = torch.cuda.Stream(self.device)
        self.depth = depth
        self._slots = None
        self.h2d_bytes_per_view = 0
        self.d2h_bytes_per_view = 0

    def _alloc(self, imgs, proj, dv):
        H, W = imgs.shape[-2:]
        h, w = H // 4, W // 4
        B = imgs.shape[0]
        slots = []
        for _ in range(self.depth):
            s = {
                "h_imgs": torch.empty(imgs.shape, dtype=imgs.dtype).pin_memory(),
                "h_proj": torch.empty(proj.shape, dtype=torch.float32).pin_memory(),
                "h_dv": torch.empty(dv.shape, dtype=torch.float32).pin_memory(),
                "d_imgs": torch.empty(imgs.shape, dtype=imgs.dtype, device=self.device),
                "d_proj": torch.empty(proj.shape, dtype=torch.float32, device=self.device),
                "d_dv": torch.empty(dv.shape, dtype=torch.float32, device=self.device),
                "h_out": torch.empty((2, B, h, w), dtype=torch.float32).pin_memory(),
                "ready": torch.cuda.Event(), "done": torch.cuda.Event(), "copied": torch.cuda.Event(),
                "graph": None, "d_out": None, "uses": 0,
            }
            slots.append(s)
        self._slots = slots
        self._shape = (tuple(imgs.shape), tuple(proj.shape), tuple(dv.shape), imgs.dtype)
        self.h2d_bytes_per_view = imgs.numel() * imgs.element_size() + 4 * (proj.numel() + dv.numel())
        self.d2h_bytes_per_view = 4 * 2 * B * h * w

    @torch.no_grad()
    def run_views(self, views, sink=None):
        """views: iterable of (imgs [B,V,3,H,W], proj [B,V,4,4], depth_values [B,D]) HOST arrays
        (numpy or CPU tensors).  imgs are float32 in [0,1] like the reference's loader output, or uint8 as decoded
        from disk (then /255 runs on the device and the upload is 4x smaller).  For every view calls sink(index, depth_np, conf_np) (numpy views of a
        pinned buffer, valid until the next call) or, without a sink, returns the list of copies."""
        results = [] if sink is None else None
        caller = torch.cuda.current_stream(self.device)
        n_streams = self.n_streams
        if n_streams == "auto":
            n_streams = 1
            it = iter(views)          # peek at the first view without materialising a lazy sweep
            first = next(it, None)
            views = [] if first is None else itertools.chain([first], it)
            if first is not None:
                shp = tuple(np.shape(first[0]))
                n_streams = 2 if int(np.prod(shp[:2] + shp[3:])) >= 2_000_000 else 1
        if n_streams == 1:
            cstreams = [caller]
        else:
            if self._compute_streams is None or len(self._compute_streams) != n_streams:
                self._compute_streams = [torch.cuda.Stream(self.device) for _ in range(n_streams)]
            cstreams = self._compute_streams
            for cs in cstreams:
                cs.wait_stream(caller)   # work the caller enqueued before this call (weight updates, ...) comes first
        pending = []  # (index, slot)

        def drain(entry):
            idx, s = entry
            s["copied"].synchronize()
            d, c = s["h_out"][0].numpy(), s["h_out"][1].numpy()
            if sink is None:
                results.append((d.copy(), c.copy()))
            else:
                sink(idx, d, c)

        for idx, (imgs, proj, dv) in enumerate(views):
            imgs = torch.as_tensor(imgs)
            if imgs.dtype != torch.uint8:
                imgs = imgs.to(torch.float32)
            proj, dv = (torch.as_tensor(a, dtype=torch.float32) for a in (proj, dv))
            if self._slots is None or self._shape != (tuple(imgs.shape), tuple(proj.shape), tuple(dv.shape), imgs.dtype):
                for e in pending:
                    drain(e)
                pending = []
                self._alloc(imgs, proj, dv)
            if len(pending) == self.depth:
                drain(pending.pop(0))
            s = self._slots[idx % self.depth]
            src = []
            for a, key in ((imgs, "h_imgs"), (proj, "h_proj"), (dv, "h_dv")):
                if a.is_pinned() and a.is_contiguous():
                    src.append(a)          # already page-locked: DMA straight from the caller's buffer
                else:
                    s[key].copy_(a)        # pageable input: stage through the slot's pinned buffer
                    src.append(s[key])
            with torch.cuda.stream(self.copy_stream):
                s["d_imgs"].copy_(src[0], non_blocking=True)
                s["d_proj"].copy_(src[1], non_blocking=True)
                s["d_dv"].copy_(src[2], non_blocking=True)
                s["ready"].record(self.copy_stream)
            compute = cstreams[idx % len(cstreams)]
            compute.wait_event(s["ready"])
            with torch.cuda.stream(compute):
                d_out = self._forward(s, compute)
            s["done"].record(compute)
            with torch.cuda.stream(self.d2h_stream):
                self.d2h_stream.wait_event(s["done"])
                s["h_out"].copy_(d_out, non_blocking=True)
                if s["graph"] is None:
                    d_out.record_stream(self.d2h_stream)
                s["copied"].record(self.d2h_stream)
            pending.append((idx, s))
        for e in pending:
            drain(e)
        if len(cstreams) > 1:
            for cs in cstreams:
                caller.wait_stream(cs)
        return results

    def _forward(self, s, compute):
        """depth + confidence [2,B,h,w] of the slot's device inputs: eagerly, or (graphs=True) by replaying the slot's
        CUDA graph.  The first two uses of a slot run eagerly (warm-up: weight packing, workspaces), the third captures."""
        if not self.graphs:
            out = self.model(s["d_imgs"], s["d_proj"], s["d_dv"])
            return torch.stack((out["depth"], out["photometric_confidence"]))
        # a captured graph holds the addresses of the packed weights: weights changed since -> capture again
        wkey = (id(self.model.feature.native_prepared()), id(self.model.cost_regularization.folded_prepared()))
        if s["graph"] is not None and s.get("wkey") != wkey:
            compute.synchronize()
            for t in self._slots:                # every graph of the pool goes, with the workspaces captured in it
                t["graph"], t["d_out"], t["uses"] = None, None, 0
            from . import ops
            ops.release_workspaces(self._capture_stream)
            self._graph_pool = torch.cuda.graph_pool_handle()
        if s["graph"] is not None:
            compute.wait_event(s["copied"])      # the previous result of this slot has left the static output tensor
            s["graph"].replay()
            return s["d_out"]
        s["uses"] += 1
        if s["uses"] < 3:                        # two eager passes: weight packing, workspaces, cache entries settled
            out = self.model(s["d_imgs"], s["d_proj"], s["d_dv"])
            return torch.stack((out["depth"], out["photometric_confidence"]))
        if self._capture_stream is None:
            self._capture_stream = torch.cuda.Stream(self.device)   # one stream for every capture: shared workspaces
            self._graph_pool = torch.cuda.graph_pool_handle()
        compute.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=self._graph_pool, stream=self._capture_stream):
            out = self.model(s["d_imgs"], s["d_proj"], s["d_dv"])
            s["d_out"] = torch.stack((out["depth"], out["photometric_confidence"]))
        s["graph"], s["wkey"] = g, wkey
        g.replay()
        return s["d_out"]

    def infer_host(self, imgs, proj, dv):
        """One reference view, host in / host out: (depth [B,h,w], confidence [B,h,w]) numpy arrays."""
        return self.run_views([(imgs, proj, dv)])[0]


def plan_scan(pairs, capacity):
    """Host-side schedule of a scan sweep (pure Python, no GPU): `pairs` is the ordered list of
    (reference image id, [source image ids]) -- the reference's pair.txt (datasets/dataloader_eval.py:41-49) -- and
    `capacity` the number of images whose features fit the device pool.  Returns one dict per reference view:
        {"ref": k, "views": [image ids, reference first], "slots": [pool slot of every view],
         "load": [(image id, pool slot), ...]}       # images whose features must be (re)computed before this view
    Eviction is least-recently-used and never touches a view of the current step."""
    slot_of, last_use, free = {}, {}, list(range(capacity - 1, -1, -1))
    steps = []
    for k, (ref, srcs) in enumerate(pairs):
        views = [int(ref)] + [int(s) for s in srcs]
        if len(set(views)) != len(views):
            raise ValueError("reference view %d: duplicate image ids %s" % (k, views))
        if len(views) > capacity:
            raise ValueError("reference view %d needs %d images, the pool holds %d" % (k, len(views), capacity))
        load = []
        for img in views:
            if img not in slot_of:
                if not free:
                    victim = min((i for i in slot_of if i not in views), key=lambda i: last_use[i])
                    free.append(slot_of.pop(victim))
                slot_of[img] = free.pop()
                load.append((img, slot_of[img]))
            last_use[img] = k
        steps.append({"ref": k, "views": views, "slots": [slot_of[i] for i in views], "load": load})
    return steps


class ScanRunner:
    """Scan-level inference with HOST buffers: the reference's eval loop (eval.py:326-360) over all reference views of
    one scan, with every image uploaded ONCE and pushed through FeatureNet ONCE per scan.  The reference re-reads and
    re-extracts the 4 source images of every reference view (dataloader_eval.py:101-176, mvsnet.py:125); in a DTU scan
    (49 images, 49 reference views x 5 views) that is 5x the uploads and 5x the FeatureNet work.  Here the fp16
    features stay in a device pool and the fused warp kernel reads the views of a depth map through an index table.

    Pipelining: image uploads run on a copy stream into a ring of staging buffers, several images ahead of the compute
    stream; depth / confidence maps return through pinned slots on a third stream, as in DepthMapRunner."""

    def __init__(self, model, device="cuda:0", pool_images=64, ring=6, depth=3, streams="auto"):
        """streams: compute streams the reference views alternate over ("auto": two from 1 M pixels per image on, see
        DepthMapRunner).  Features written on one stream are handed to the other through events, and a pool slot is
        only refilled after every forward pass that read it -- on either stream -- has finished."""
        self.model = model.to(device).eval()
        self.device = torch.device(device)
        self.pool_images, self.ring, self.depth = pool_images, ring, depth
        self.n_streams = streams if streams == "auto" else max(1, int(streams))
        self._compute_streams = None
        self.copy_stream = torch.cuda.Stream(self.device)
        self.d2h_stream = torch.cuda.Stream(self.device)
        self._key = None
        self.h2d_bytes = 0   # of the last run_scan
        self.d2h_bytes = 0
        self.featurenet_images = 0

    def _alloc(self, img_shape, dtype, n_out):
        C, H, W = img_shape
        h, w = H // 4, W // 4
        dev = self.device
        self.pool = torch.empty((self.pool_images, h, 4, w, 8), dtype=torch.float16, device=dev)
        self.stage = [{"d": torch.empty((C, H, W), dtype=dtype, device=dev), "h": None,
                       "ready": torch.cuda.Event(), "free": torch.cuda.Event(), "used": False} for _ in range(self.ring)]
        self.outs = [{"h": torch.empty((2, 1, h, w), dtype=torch.float32).pin_memory(), "done": torch.cuda.Event(),
                      "copied": torch.cuda.Event()} for _ in range(n_out)]
        self._key = (tuple(img_shape), dtype)

    @torch.no_grad()
    def run_scan(self, images, projs, depth_values, pairs, sink=None):
        """images: sequence of HOST arrays [3,H,W] (float32 in [0,1] like the reference loader's output, or uint8 as
        decoded from disk); projs [n_images,4,4]; depth_values [D] (shared) or [n_ref, D]; pairs: ordered
        [(ref image id, [source image ids])].  Calls sink(k, depth_np, conf_np) per reference view k (numpy views of a
        pinned buffer, valid until the next call) or, without a sink, returns the list of (depth, conf) copies."""
        steps = plan_scan(pairs, self.pool_images)
        first = torch.as_tensor(images[0])
        dtype = torch.uint8 if first.dtype == torch.uint8 else torch.float32
        if self._key != (tuple(first.shape), dtype):
            self._alloc(tuple(first.shape), dtype, self.depth)
        dev, caller = self.device, torch.cuda.current_stream(self.device)
        n_streams = self.n_streams
        if n_streams == "auto":
            n_streams = 2 if int(first.shape[-1]) * int(first.shape[-2]) >= 1_000_000 else 1
        if n_streams == 1:
            cstreams = [caller]
        else:
            if self._compute_streams is None or len(self._compute_streams) != n_streams:
                self._compute_streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
            cstreams = self._compute_streams
        projs = torch.as_tensor(projs, dtype=torch.float32)
        dvs = torch.as_tensor(depth_values, dtype=torch.float32)
        if dvs.dim() == 1:
            dvs = dvs.unsqueeze(0)
        # per-step projection matrices, assembled on the host and uploaded once (a few KB)
        flat = torch.cat([projs[s["views"]] for s in steps]).contiguous()
        d_proj = flat.to(dev, non_blocking=False)
        d_dv = dvs.contiguous().to(dev)
        self.h2d_bytes = 4 * (flat.numel() + dvs.numel())
        self.d2h_bytes = 0
        self.featurenet_images = 0

        for cs in cstreams:
            if cs is not caller:
                cs.wait_stream(caller)       # the small uploads above, and whatever the caller enqueued before
        uploads = [(k, img, slot) for k, s in enumerate(steps) for img, slot in s["load"]]
        issued = consumed = 0
        results = [] if sink is None else None
        pending = []
        slot_written = {}   # pool slot -> (event after the FeatureNet pass that filled it, index of its stream)
        slot_read = {}      # pool slot -> {stream index: event after the last forward pass on that stream that read it}

        def issue(u):
            _, img, _ = uploads[u]
            st = self.stage[u % self.ring]
            src = torch.as_tensor(images[img])
            if src.dtype != dtype:
                src = src.to(dtype)
            if not (src.is_pinned() and src.is_contiguous()):
                if st["h"] is None:
                    st["h"] = torch.empty(src.shape, dtype=dtype).pin_memory()
                if st["used"]:
                    st["ready"].synchronize()   # the previous upload out of this pinned buffer has finished
                st["h"].copy_(src)
                src = st["h"]
            with torch.cuda.stream(self.copy_stream):
                if st["used"]:
                    self.copy_stream.wait_event(st["free"])  # FeatureNet has consumed the previous occupant
                st["d"].copy_(src, non_blocking=True)
                st["ready"].record(self.copy_stream)
            st["used"] = True
            self.h2d_bytes += src.numel() * src.element_size()

        def drain(entry):
            k, o = entry
            o["copied"].synchronize()
            d, c = o["h"][0].numpy(), o["h"][1].numpy()
            if sink is None:
                results.append((d.copy(), c.copy()))
            else:
                sink(k, d, c)

        off = 0
        for k, s in enumerate(steps):
            ci = k % len(cstreams)
            compute = cstreams[ci]
            # keep the copy engine ahead of the compute streams: as many uploads in flight as the ring holds
            while issued < len(uploads) and issued - consumed < self.ring:
                issue(issued)
                issued += 1
            for img, slot in s["load"]:
                if consumed == issued:
                    issue(issued)
                    issued += 1
                st = self.stage[consumed % self.ring]
                compute.wait_event(st["ready"])
                for cj, ev in slot_read.pop(slot, {}).items():   # the slot's previous occupant is no longer read
                    if cj != ci:
                        compute.wait_event(ev)
                with torch.cuda.stream(compute):
                    self.model.features_to_pool(st["d"].unsqueeze(0), self.pool[slot:slot + 1])
                st["free"].record(compute)
                if len(cstreams) > 1:
                    ev = torch.cuda.Event()
                    ev.record(compute)
                    slot_written[slot] = (ev, ci)
                consumed += 1
                self.featurenet_images += 1
                while issued < len(uploads) and issued - consumed < self.ring:
                    issue(issued)
                    issued += 1
            V = len(s["views"])
            if len(cstreams) > 1:
                for slot in s["slots"]:                          # features filled on the other stream
                    ev, cj = slot_written[slot]
                    if cj != ci:
                        compute.wait_event(ev)
            with torch.cuda.stream(compute):
                out = self.model.forward_from_pool(self.pool, s["slots"], d_proj[off:off + V].unsqueeze(0),
                                                   d_dv[min(k, d_dv.shape[0] - 1)].unsqueeze(0))
                d_out = torch.stack((out["depth"], out["photometric_confidence"]))
            if len(cstreams) > 1:
                rev = torch.cuda.Event()
                rev.record(compute)
                for slot in s["slots"]:
                    slot_read.setdefault(slot, {})[ci] = rev
            off += V
            if len(pending) == self.depth:
                drain(pending.pop(0))
            o = self.outs[k % self.depth]
            o["done"].record(compute)
            with torch.cuda.stream(self.d2h_stream):
                self.d2h_stream.wait_event(o["done"])
                o["h"].copy_(d_out, non_blocking=True)
                d_out.record_stream(self.d2h_stream)
                o["copied"].record(self.d2h_stream)
            self.d2h_bytes += d_out.numel() * 4
            pending.append((k, o))
        for e in pending:
            drain(e)
        for st in self.stage:
            st["used"] = False
        for cs in cstreams:
            if cs is not caller:
                caller.wait_stream(cs)
        caller.synchronize()
        return results
