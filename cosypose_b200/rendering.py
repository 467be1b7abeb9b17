"""Pre-rendered views standing in for the renderer side input.

The path treats the rendered view as an input (reference renderer contract:
cosypose/models/pose.py:100-102 calling cosypose/rendering/bullet_batch_renderer.py:46-90, which
returns float [B,3,240,320] in [0,1]).  `PreRenderedViews` replays stacks generated up front, in
the order `CoarseRefinePosePredictor` asks for them: per stage, per chunk of `bsz_objects`
hypotheses, per iteration.
"""
import torch


class PreRenderedViews:
    def __init__(self, stages, bsz_objects=64, device=None):
        """stages: list of tensors [n_iter_s, N, 3, 240, 320], one per stage in call order
        (e.g. [coarse_views, refiner_views])."""
        self.bsz = bsz_objects
        self.chunks = []      # flat list of [n_iter, Bc, 3, H, W] contiguous tensors, call order
        for views in stages:
            n = views.shape[1]
            for s in range(0, n, bsz_objects):
                c = views[:, s:s + bsz_objects].contiguous()
                self.chunks.append(c.to(device) if device is not None else c)
        self.reset()

    @classmethod
    def from_uint8(cls, stages, bsz_objects=64):
        """stages: device tensors [n_iter_s, N, 240, 320, 3] uint8 (NHWC, as the reference's renderer
        delivers them); consumed by the engine without a conversion pass."""
        self = cls.__new__(cls)
        self.bsz = bsz_objects
        self.chunks = []
        for views in stages:
            for s in range(0, views.shape[1], bsz_objects):
                self.chunks.append(views[:, s:s + bsz_objects].contiguous())
        self.reset()
        return self

    @classmethod
    def from_chunks(cls, chunks, bsz_objects=64):
        """chunks: per (stage, chunk) stacks [n_iter, Bc, ...] in call order, used in place (no copy): a caller
        that refills the same device buffers every step keeps the engine's captured graphs valid."""
        self = cls.__new__(cls)
        self.bsz = bsz_objects
        self.chunks = [c for c in chunks]
        assert all(c.is_contiguous() for c in self.chunks)
        self.reset()
        return self

    def reset(self):
        self._chunk = 0
        self._it = 0

    def _advance(self, n_iter_taken):
        self._it += n_iter_taken
        if self._it >= self.chunks[self._chunk].shape[0]:
            self._it = 0
            self._chunk = (self._chunk + 1) % len(self.chunks)

    def prerendered(self, n_iterations, batch_size):
        """All views of the next `n_iterations` calls as one [n_iterations, B, 3, H, W] stack."""
        c = self.chunks[self._chunk]
        assert self._it == 0 and c.shape[0] == n_iterations and c.shape[1] == batch_size, \
            'pre-rendered stack does not match the call sequence'
        self._advance(n_iterations)
        return c

    def render(self, obj_infos, TCO, K, resolution=(240, 320), **kwargs):
        c = self.chunks[self._chunk]
        assert c.shape[1] == len(obj_infos), 'pre-rendered stack does not match the call sequence'
        out = c[self._it]
        self._advance(1)
        return out


class PerCallRenderer:
    """Wraps PreRenderedViews but hides `prerendered`, forcing the two-phase per-iteration path
    (prepare_iter -> render -> refine_iter) a real renderer needs."""

    def __init__(self, views):
        self.views = views

    def reset(self):
        self.views.reset()

    def render(self, *args, **kwargs):
        return self.views.render(*args, **kwargs)
