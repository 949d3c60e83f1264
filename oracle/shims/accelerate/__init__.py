"""Stand-in for HuggingFace `accelerate` so that the reference module imports (VDDP:24, main.py:5).
No behaviour: golden generation never builds a Trainer.  Test infrastructure only."""


class Accelerator:  # pragma: no cover - placeholder
    pass


class DistributedDataParallelKwargs:  # pragma: no cover
    def __init__(self, **kw):
        self.kw = kw


class InitProcessGroupKwargs:  # pragma: no cover
    def __init__(self, **kw):
        self.kw = kw
