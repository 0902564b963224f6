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


if __name__ == "__main__":
    main()
