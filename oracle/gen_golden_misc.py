"""oracle/gen_golden_misc.py -- TEST INFRASTRUCTURE, build container only (needs /root/reference).

    python -m oracle.gen_golden_misc      # writes tests/golden/temporal_agg.npz

`src/utils/misc.py` imports clip / lightning at module level and cannot be imported here; the source text of
`class TemporalAgg` (misc.py:88-140) is cut out of the reference file and executed unmodified against numpy."""
import ast
from pathlib import Path

import numpy as np

REF = Path("/root/reference/src/utils/misc.py")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "temporal_agg.npz"


def reference_class():
    src = REF.read_text()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "TemporalAgg")
    ns = {"np": np}
    exec(compile(ast.Module(body=[node], type_ignores=[]), str(REF), "exec"), ns)
    return ns["TemporalAgg"]


def rotation_vectors(out):
    """`rotation_6d_to_matrix` + `matrix_to_quaternion` (src/utils/rotation_conversions.py:102-161,556-577) -- the
    rot6d -> quaternion conversion of ACTRLBenchPCD's inference branch (act.py:785-795); the file is pure torch and is
    loaded by path."""
    import importlib.util

    import torch

    spec = importlib.util.spec_from_file_location("ref_rotation_conversions", "/root/reference/src/utils/rotation_conversions.py")
    rc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rc)
    g = torch.Generator().manual_seed(12)
    d6 = torch.randn(64, 5, 6, generator=g)
    d6[0, 0] = torch.tensor([1.0, 0, 0, 0, 1.0, 0])          # identity
    d6[0, 1] = torch.tensor([-1.0, 0, 0, 0, -1.0, 0])        # 180 degrees about z: w = 0 branch
    d6[0, 2] = torch.tensor([1.0, 0, 0, 0, -1.0, 0])         # 180 degrees about x
    m = rc.rotation_6d_to_matrix(d6)
    out["rot/d6"], out["rot/matrix"], out["rot/quat"] = d6.numpy(), m.numpy(), rc.matrix_to_quaternion(m).numpy()


def normalizer_vectors(out):
    """`_fit` / `_normalize` of src/utils/diffusion_policy/normalizer.py:195-300 (limits and gaussian modes, a constant
    channel exercising range_eps), run through the reference's own LinearNormalizer."""
    import torch

    from .gen_golden_dp import install_dp_shim

    install_dp_shim()
    import sys

    LN = sys.modules["src.utils.diffusion_policy"].LinearNormalizer
    g = torch.Generator().manual_seed(3)
    data = torch.randn(200, 16, 7, generator=g) * torch.tensor([1.0, 5.0, 0.1, 2.0, 1.0, 0.01, 3.0]) + torch.arange(7.0)
    data[..., 4] = 2.5  # constant channel
    x = torch.randn(9, 16, 7, generator=g)
    out["norm/data"], out["norm/x"] = data.numpy(), x.numpy()
    for mode, kw in (("limits", {}), ("gaussian", {}), ("limits_nooffset", {"fit_offset": False})):
        n = LN()
        n.fit({"action": data}, last_n_dims=1, mode=mode.split("_")[0], **kw)
        out[f"norm/{mode}/scale"] = n.params_dict["action"]["scale"].detach().numpy()
        out[f"norm/{mode}/offset"] = n.params_dict["action"]["offset"].detach().numpy()
        out[f"norm/{mode}/y"] = n["action"].normalize(x).detach().numpy()
        out[f"norm/{mode}/back"] = n["action"].unnormalize(n["action"].normalize(x)).detach().numpy()


def main():
    cls = reference_class()
    rng = np.random.default_rng(5)
    out = {}
    for name, (chunk, dim, k, steps) in {"a": (6, 3, 0.01, 15), "b": (4, 7, 0.25, 9)}.items():
        agg = cls(apply=True, action_dim=dim, chunk_size=chunk, k=k)
        chunks = rng.normal(size=(steps, chunk, dim))
        out[f"{name}/cfg"] = np.array([chunk, dim, k, steps])
        out[f"{name}/chunks"] = chunks
        out[f"{name}/actions"] = np.stack([agg(c) for c in chunks])
    np.savez_compressed(OUT, **out)
    print(OUT)
    extra = {}
    rotation_vectors(extra)
    normalizer_vectors(extra)
    path = OUT.parent / "misc_rotation_normalizer.npz"
    np.savez_compressed(path, **extra)
    print(path)


if __name__ == "__main__":
    main()
