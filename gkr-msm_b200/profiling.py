"""Optional span accounting for the host-side protocol mirror (the reference prints a tracing span tree,
examples/pippenger.rs:75-89).  Enabled by setting `profiling.PROFILE = {}`; every span synchronises the context on both
sides, so enable it only to look at a breakdown, never for a timed run."""
from __future__ import annotations

import time

PROFILE = None  # name -> seconds


class span:
    def __init__(self, ctx, name):
        self.ctx, self.name = ctx, name

    def __enter__(self):
        if PROFILE is not None:
            self.ctx.sync()
            self.t0 = time.perf_counter()
        return self

    def __exit__(self, *a):
        if PROFILE is not None:
            self.ctx.sync()
            PROFILE[self.name] = PROFILE.get(self.name, 0.0) + time.perf_counter() - self.t0
