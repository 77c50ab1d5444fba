"""vae_b200 -- B200-native Variational Factorization Machine step.

Drop-in replacements for the model classes of the reference's two training
scripts, backed by hand-written sm_100a CUDA kernels behind a C ABI
(``include/vfm_b200.h``):

* ``vae_b200.vfm_torch.CF``     <- ``vfm-torch.py``   (sampled ELBO)
* ``vae_b200.vfm_tomasrch.CF``  <- ``vfm-tomasrch.py`` (closed-form Gaussian)

No CPU fallback: importing works anywhere, computing requires a CUDA device
and the built ``libvfm_b200.so``.
"""
__version__ = "0.1.0"

__all__ = ["SampledCF", "ClosedCF", "build_extension"]


def build_extension(force: bool = False, verbose: bool = False) -> str:
    from . import _lib
    return _lib.build(force=force, verbose=verbose)


def __getattr__(name):
    if name == "SampledCF":
        from .vfm_torch import CF
        return CF
    if name == "ClosedCF":
        from .vfm_tomasrch import CF
        return CF
    raise AttributeError(name)
