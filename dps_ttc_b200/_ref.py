"""Locating and importing the reference tree (vishnutez/dps-ttc).

The reference is NOT part of this package.  It is needed for exactly two things:
  * the UNet ε-predictor (`guided_diffusion.unet.create_model`), which by the task's north star stays
    the reference's own PyTorch module — "the model, not the graft";
  * re-binding names in the reference's registries so its unchanged drivers pick the B200 classes
    (`install_into_reference`), and, in tests / the bench's CPU arm, running the reference itself.
Search order: $DPS_REF, /root/reference (this container), <repo>/baseline/_ref (staged copy that
travels to the GPU box; created by __graft_entry__.build(), git-ignored).
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
STUB_DIR = os.path.join(_HERE, "_stubs")
STAGED = os.path.join(REPO_ROOT, "baseline", "_ref")


def reference_root():
    for cand in (os.environ.get("DPS_REF"), "/root/reference", STAGED):
        if cand and os.path.isfile(os.path.join(cand, "guided_diffusion", "unet.py")):
            return cand
    return None


def _missing(mod: str) -> bool:
    try:
        return importlib.util.find_spec(mod) is None
    except (ImportError, ValueError):
        return True


def ensure_reference(required: bool = True):
    """Put the reference (and stubs for packages that are really absent) on sys.path."""
    root = reference_root()
    if root is None:
        if required:
            raise RuntimeError("reference tree not found (set DPS_REF, or run __graft_entry__.build() where "
                               "/root/reference exists to stage baseline/_ref)")
        return None
    sys.dont_write_bytecode = True  # /root/reference is read-only
    if any(_missing(mod) for mod in ("matplotlib", "motionblur", "facenet_pytorch")):
        if STUB_DIR not in sys.path:
            sys.path.append(STUB_DIR)  # appended: a really installed package always wins
    if root not in sys.path:
        sys.path.insert(0, root)
    return root


@contextlib.contextmanager
def quiet():
    """The reference prints from inside its hot loop; silence it where we call it."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def load_yaml(path):
    import yaml
    with open(path) as f:
        return yaml.load(f, Loader=yaml.FullLoader)


def reference_config(name: str):
    root = ensure_reference()
    return load_yaml(os.path.join(root, "configs", name))


def create_unet(config: str = "model_config.yaml", reinit_zero_seed=None, device="cpu"):
    """The reference's UNetModel with random-init weights (the checkpoint path in the YAML does not
    exist offline; create_model swallows the failure, unet.py:87-90).

    Random init zeroes every output conv (zero_module, nn.py:68-74), so ε ≡ 0 and the VJP ≡ 0
    (SURVEY §0).  `reinit_zero_seed` re-draws all-zero weight tensors from N(0, 0.02²) under a fixed
    seed so that the ε / VJP data path is exercised; the same module object serves both sides of every
    comparison, so parity is unaffected."""
    import torch
    ensure_reference()
    cfg = dict(reference_config(config))
    with quiet():
        from guided_diffusion.unet import create_model
        model = create_model(**cfg)
    if reinit_zero_seed is not None:
        gen = torch.Generator().manual_seed(int(reinit_zero_seed))
        with torch.no_grad():
            for p in model.parameters():
                if p.ndim > 1 and not p.any():
                    p.copy_(torch.randn(p.shape, generator=gen) * 0.02)
    return model.to(device).eval()


def install_into_reference(suffix: str = "", rebind: bool = True):
    """Make the B200 classes selectable from the reference's YAML/drivers without editing them.

    rebind=True  : assign into the reference's registry dicts under the SAME names
                   (`__OPERATOR__['gaussian_blur'] = B200 class`, …), the documented way to replace an
                   entry since register_* raises on duplicates (measurements.py:22-23);
    suffix='_b200': additionally/alternatively register under new names to select in YAML."""
    ensure_reference()
    from . import conditioning, operators, sampler  # noqa: F401  (populate our registries)
    from .registry import CONDITIONING, NOISES, OPERATORS, SAMPLERS
    with quiet():
        import guided_diffusion.condition_methods as ref_cond
        import guided_diffusion.gaussian_diffusion as ref_gd
        import guided_diffusion.measurements as ref_meas
    pairs = ((ref_meas.__OPERATOR__, OPERATORS), (ref_meas.__NOISE__, NOISES),
             (ref_cond.__CONDITIONING_METHOD__, CONDITIONING), (ref_gd.__SAMPLER__, SAMPLERS))
    for ref_table, ours in pairs:
        for name, cls in ours.table.items():
            if rebind:
                ref_table[name] = cls
            if suffix:
                ref_table[name + suffix] = cls
    return ref_meas, ref_cond, ref_gd
