"""Regenerates tests/golden/ from the reference tree (run in the build container only; /root/reference
does not exist on the GPU box).  Fixtures are DATA samples the reference ships, not source code.

  AL_12B_grid_128r.rle      verbatim copy of docs/decompress/samples/AL_12B_grid_128r.rle (257 874 B)
  AL_12B_grid_128r.npz      the shipped decoder output docs/decompress/samples/AL_12B_grid_128r.npy
                            (raw uint8 128^3 written by decompress_grid.py:35 with ndarray.tofile), compressed
  golden.json               sha256 of both + dims / occupancy facts quoted in SURVEY.md finding 10
"""
import hashlib, json, os, shutil, struct
import numpy as np

REF = "/root/reference/docs/decompress/samples"
HERE = os.path.dirname(os.path.abspath(__file__))

rle = open(f"{REF}/AL_12B_grid_128r.rle", "rb").read()
npy = np.fromfile(f"{REF}/AL_12B_grid_128r.npy", dtype=np.uint8)
shutil.copyfile(f"{REF}/AL_12B_grid_128r.rle", f"{HERE}/AL_12B_grid_128r.rle")
os.chmod(f"{HERE}/AL_12B_grid_128r.rle", 0o644)
np.savez_compressed(f"{HERE}/AL_12B_grid_128r.npz", grid=npy.reshape(128, 128, 128))
meta = {
    "rle_sha256": hashlib.sha256(rle).hexdigest(),
    "npy_sha256": hashlib.sha256(npy.tobytes()).hexdigest(),
    "dims": list(struct.unpack("<III", rle[:12])),
    "runs": (len(rle) - 12) // 6,
    "occupied": int(npy.sum()),
    "source": "AlfonsoLRz/VoxelFragmentML docs/decompress/samples/ (reference fixture)",
}
json.dump(meta, open(f"{HERE}/golden.json", "w"), indent=1)
print(meta)
