"""Tiny CUDA-graph cache: capture a fixed-shape launch sequence once, replay it with fresh inputs copied into static buffers."""
from __future__ import annotations

import torch


class GraphCache:
    def __init__(self):
        self._entries = {}

    def clear(self):
        self._entries.clear()

    def run(self, key, fn, *inputs):
        """fn(*static_inputs) -> tensor (or tuple of tensors) built only from stream-ordered work; inputs must keep shape/dtype per key."""
        ent = self._entries.get(key)
        if ent is None:
            static_in = [t.clone() for t in inputs]
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):  # warm-up: lazy initialisation (function attributes, weight packing) must not be captured
                fn(*static_in)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_out = fn(*static_in)
            ent = (g, static_in, static_out)
            self._entries[key] = ent
        g, static_in, static_out = ent
        for dst, src in zip(static_in, inputs):
            dst.copy_(src)
        g.replay()
        if isinstance(static_out, (tuple, list)):
            return type(static_out)(o.clone() for o in static_out)
        return static_out.clone()
