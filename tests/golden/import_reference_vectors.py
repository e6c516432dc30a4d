"""Pack the output directory of julia/emit_golden.jl (raw column-major .bin files + manifest.json, written by the REFERENCE
under a Julia runtime) into tests/golden/reference_vectors.npz, the file tests/test_reference_vectors.py looks for.

    julia --project=/path/to/Particulator.jl julia/emit_golden.jl /tmp/refvec
    python tests/golden/import_reference_vectors.py /tmp/refvec"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DTYPES = {"Float64": np.float64, "Int64": np.int64, "UInt8": np.uint8}


def pack(src, dst=os.path.join(HERE, "reference_vectors.npz")):
    man = json.load(open(os.path.join(src, "manifest.json")))
    if man.get("format") != "particulator_b200.reference_vectors":
        raise ValueError("not an emit_golden.jl output directory")
    arrays = {}
    for name, meta in man["arrays"].items():
        raw = np.fromfile(os.path.join(src, name + ".bin"), dtype=DTYPES[meta["dtype"]])
        arrays[name] = raw.reshape(meta["shape"], order="F")
    arrays["__julia_version__"] = np.frombuffer(str(man.get("julia", "?")).encode(), dtype=np.uint8)
    np.savez_compressed(dst, **arrays)
    return dst, len(arrays)


if __name__ == "__main__":
    print(*pack(sys.argv[1], *(sys.argv[2:3])))
