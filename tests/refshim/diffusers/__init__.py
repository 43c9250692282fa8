"""Import shim so the reference's sampler modules import without the real `diffusers` (not installed, no
network).  Only the four symbols the reference imports at module scope are provided; none is used at run time by
the paths the pin tests exercise (the oracle supplies the model object)."""


class StableDiffusionPipeline:  # pragma: no cover - placeholder
    pass


class DDIMScheduler:  # pragma: no cover - placeholder
    pass
