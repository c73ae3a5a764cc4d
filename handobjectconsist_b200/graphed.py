"""CUDA-graph capture of the consistency step.

``warpbranch.forward`` + its backward is ~60 launches of which ~25 are this package's kernels and the rest
small tensor ops; at 16 pairs of 256x256 the GPU work is well under a millisecond, so eager execution is
bound by CPU launch overhead.  ``GraphedConsistStep`` captures forward AND backward (loss -> gradients of the
first frame's hand / object vertices) once for fixed shapes and replays the graph; inputs are copied into
static device buffers (straight from pinned host memory when they live there), outputs are read from
static buffers.  ``GraphedConsistStep.apply`` is the autograd-friendly entry: it returns the loss as a
differentiable function of the predicted vertices, so it drops into a training step in place of
``WarpRegNet.warp_forward`` (/root/reference/meshreg/models/warpreg.py:66-79).
"""
import torch
from torch.autograd import Function

from . import warpbranch


def _name(key):
    return getattr(key, "name", key)


class GraphedConsistStep:
    def __init__(self, renderer, criterion, image_size, hand_face, samples, all_results, hand_ignore_faces=None,
                 gt_refs=True, first_only=True, use_backward=True, detach_renders=True, warmup=3,
                 before_capture=None, return_visuals=False):
        """``samples`` / ``all_results``: one example batch (reference layout, warpbranch.py:27-44) that fixes
        shapes and dtypes; their values are only used for the warm-up iterations."""
        if len(samples) != 2:
            raise ValueError("GraphedConsistStep captures one frame pair (sample_nb == 2)")
        self.device = torch.device("cuda", torch.cuda.current_device())
        dev = self.device
        self.renderer, self.criterion, self.image_size = renderer, criterion, image_size
        # the captured step returns the loss and its gradients only: the visualisation entries of pair_results
        # (warps / diffs / warp_mask) would be dead stores inside the graph (return_visuals=True keeps them, for
        # benchmarking the difference)
        self.kw = dict(gt_refs=gt_refs, first_only=first_only, hand_ignore_faces=hand_ignore_faces,
                       use_backward=use_backward, detach_renders=detach_renders, return_visuals=return_visuals,
                       loss_only=not return_visuals)  # (the captured step hands out the loss and its gradients only)
        self.hand_face = hand_face.to(dev)
        self._u8_stage = {}
        self._one = torch.ones((), dtype=torch.float32, device=dev)  # d loss / d loss, created once (not per replay)
        # static inputs
        def static(k, v):
            if not torch.is_tensor(v):
                return v
            if v.dtype == torch.uint8 and _name(k) in ("IMAGE", "JITTERMASK"):  # uint8 frames: see _load_u8
                f = v.to(dev).float().div(255.0)
                return f - 0.5 if _name(k) == "IMAGE" else f
            return v.to(dev).clone()

        self.samples = [{k: static(k, v) for k, v in s.items()} for s in samples]
        self.results = [{k: v.detach().to(dev).clone() for k, v in r.items() if torch.is_tensor(v)}
                        for r in all_results]
        self.hand = self.results[0]["recov_handverts3d"].requires_grad_(True)
        self.obj = self.results[0]["recov_objverts3d"].requires_grad_(True)

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if before_capture is not None:
            before_capture()  # e.g. arm the library's device timer so that the capture is instrumented
        self.graph = torch.cuda.CUDAGraph()
        # capture on a HIGH-priority stream: the chain that stays on it (the render whose geometry gradient is needed,
        # i.e. scan -> cover -> line pass of the rasterizer backward) is the critical path of the step, and kernel
        # nodes inherit the priority of the stream they were captured on -- the side stream's work (default priority)
        # fills in around it instead of delaying it
        self.capture_stream = torch.cuda.Stream(device=dev, priority=-1)
        with torch.cuda.graph(self.graph, stream=self.capture_stream):
            self._captured = self._run()
        self.loss, self.grad_hand, self.grad_obj = self._captured[:3]

    def _run(self):
        loss, _ = warpbranch.forward(self.samples, self.results, self.hand_face, self.renderer, self.image_size,
                                     self.criterion, **self.kw)
        gh, go = torch.autograd.grad(loss, [self.hand, self.obj], grad_outputs=self._one, allow_unused=True)
        if gh is None:
            gh = torch.zeros_like(self.hand)
        if go is None:
            go = torch.zeros_like(self.obj)
        return loss.detach(), gh, go

    def _load_u8(self, buf, src, name):
        """uint8 frame / jitter mask -> the static fp32 buffer: the bytes cross PCIe as uint8 (a quarter of fp32) and
        are widened on the device with the two operations of the reference's `to_tensor` + `normalize`
        (handobjset.py:368-379): x / 255 - 0.5 for IMAGE, x / 255 for JITTERMASK."""
        from . import _lib
        key = id(buf)
        stage = self._u8_stage.get(key)
        if stage is None or stage.shape != src.shape:
            stage = self._u8_stage[key] = torch.empty(src.shape, dtype=torch.uint8, device=self.device)
        sub = 0.5 if name == "IMAGE" else 0.0
        with torch.cuda.device(self.device):  # the launch goes to THIS step's device, whatever the caller's current one
            stage.copy_(src, non_blocking=True)
            _lib.check(_lib.lib().hoc_unpack_u8(_lib.ptr(stage), _lib.ptr(buf), buf.numel(), 255.0, sub,
                                                _lib.stream_ptr()), "hoc_unpack_u8")

    def load(self, samples, all_results, skip_images=False):
        """Copy a new batch into the static buffers (stream-ordered; pinned host tensors copy asynchronously).
        IMAGE / JITTERMASK may be given as uint8 tensors (0..255): see ``_load_u8``; ``skip_images=True`` leaves them
        alone (they come from ``load_frames``)."""
        with torch.no_grad():
            for dst, src in zip(self.samples, samples):
                by_name = {(_name(k), type(k).__name__): v for k, v in src.items()}
                for k, buf in dst.items():
                    if skip_images and _name(k) in ("IMAGE", "JITTERMASK"):
                        continue
                    if torch.is_tensor(buf):
                        val = by_name[(_name(k), type(k).__name__)]
                        if val.dtype == torch.uint8 and buf.dtype == torch.float32 and _name(k) in ("IMAGE", "JITTERMASK"):
                            self._load_u8(buf, val, _name(k))
                        else:
                            buf.copy_(val, non_blocking=True)
            for dst, src in zip(self.results, all_results):
                for k, buf in dst.items():
                    buf.copy_(src[k].detach(), non_blocking=True)

    def load_frames(self, frames, affinetrans, color=None, orders=None):
        """The image side of a new batch straight from the DECODED frames: two uint8 tensors [B,Hs,Ws,3] (pinned host
        memory: a quarter of the bytes of the float tensors) are colour-jittered, cropped / rotated to ``image_size``,
        normalised and masked on the device (``inputpipe.augment_frame_pair``, two launches, bit-compatible with the
        PIL calls of handobjset.py:336-379) directly into this step's static IMAGE / JITTERMASK buffers.  The other
        sample entries (intrinsics, faces, reference meshes, predicted vertices) still go through ``load``."""
        from . import inputpipe

        def buf(sample, name):
            for k, v in sample.items():
                if _name(k) == name and torch.is_tensor(v) and v.dim() == 4:
                    return v
            raise KeyError(name)

        images = [buf(s, "IMAGE") for s in self.samples]
        masks = [buf(s, "JITTERMASK") for s in self.samples]
        with torch.cuda.device(self.device):
            inputpipe.augment_frame_pair(frames, affinetrans, self.image_size, color=color, orders=orders,
                                         out=(images, masks))

    def replay(self):
        """Run the captured forward + backward; returns the static (loss, grad_hand, grad_obj) tensors."""
        self.graph.replay()
        return self.loss, self.grad_hand, self.grad_obj

    def __call__(self, samples, all_results):
        self.load(samples, all_results)
        return self.replay()

    def apply(self, samples, all_results):
        """Differentiable entry: loss as a function of all_results[0]['recov_handverts3d' / 'recov_objverts3d']."""
        return _GraphedConsistFunction.apply(all_results[0]["recov_handverts3d"], all_results[0]["recov_objverts3d"],
                                             self, samples, all_results)


class _GraphedConsistFunction(Function):
    @staticmethod
    def forward(ctx, hand, obj, step, samples, all_results):
        loss, gh, go = step(samples, all_results)
        ctx.save_for_backward(gh.clone(), go.clone())
        return loss.clone()

    @staticmethod
    def backward(ctx, grad_loss):
        gh, go = ctx.saved_tensors
        return gh * grad_loss, go * grad_loss, None, None, None


class GraphedHeadConsistStep(GraphedConsistStep):
    """The captured step with the geometry head (and whatever else ``head`` runs: MANO skinning, ManoAdaptor,
    recover_3d_proj, ObjBranch) INSIDE the graph: network outputs -> loss -> gradients of the network outputs in
    one replay, instead of a launch-bound eager front end (~1 ms of Python for ~0.1 ms of kernels) in front of it.

    ``head(inputs)`` maps a dict of device tensors (pose, shape, scale / translation / rotation heads, canonical object
    vertices ...) to ``(recov_handverts3d, recov_objverts3d)`` of the first frame -- what MeshRegNet.recover_mano /
    recover_object produce (/root/reference/meshreg/models/meshregnet.py:179-272) -- using capturable operators only
    (no host synchronisation, no host-to-device copies: everything it reads lives on the device).
    ``head_inputs`` is an example dict that fixes shapes; every floating-point entry gets a gradient
    (``head_grads[name]``), other entries are constants of the step.  ``all_results[0]`` must hold example
    ``recov_handverts3d`` / ``recov_objverts3d`` at construction (shapes); batches loaded later need not."""

    def __init__(self, head, head_inputs, renderer, criterion, image_size, hand_face, samples, all_results, **kw):
        dev = torch.device("cuda", torch.cuda.current_device())
        self.head = head
        self.head_inputs = {k: v.detach().to(dev).clone().requires_grad_(v.is_floating_point())
                            for k, v in head_inputs.items()}
        self._grad_keys = [k for k, v in self.head_inputs.items() if v.requires_grad]
        super().__init__(renderer, criterion, image_size, hand_face, samples, all_results, **kw)
        self.head_grads = dict(zip(self._grad_keys, self._captured[3:]))
        # the first frame's vertices now come from the head: later batches need not carry them
        for key in ("recov_handverts3d", "recov_objverts3d"):
            self.results[0].pop(key, None)

    def _run(self):
        hand, obj = self.head(self.head_inputs)
        results = [dict(self.results[0], recov_handverts3d=hand, recov_objverts3d=obj)] + self.results[1:]
        loss, _ = warpbranch.forward(self.samples, results, self.hand_face, self.renderer, self.image_size,
                                     self.criterion, **self.kw)
        wrt = [hand, obj] + [self.head_inputs[k] for k in self._grad_keys]
        grads = torch.autograd.grad(loss, wrt, grad_outputs=self._one, allow_unused=True)
        return (loss.detach(),) + tuple(torch.zeros_like(w) if g is None else g for g, w in zip(grads, wrt))

    def load(self, samples, all_results, head_inputs=None):
        super().load(samples, all_results)
        if head_inputs is not None:
            with torch.no_grad():
                for k, buf in self.head_inputs.items():
                    buf.copy_(head_inputs[k].detach(), non_blocking=True)

    def __call__(self, samples, all_results, head_inputs=None):
        """Returns the static ``(loss, head_grads)``; ``grad_hand`` / ``grad_obj`` (gradients that reach the mesh
        vertices) stay available as attributes."""
        self.load(samples, all_results, head_inputs)
        self.graph.replay()
        return self.loss, self.head_grads
