"""The position-weighted 64-bit hash bench.py prints as `parity.hash` (sum_j byte_j * ((j + 1) * 0x9E3779B97F4A7C15 mod
2^64) mod 2^64), of a FILE, on the CPU -- to pin a device-computed hash against an output the reference binary wrote.

    python tools/file_hash.py out.bedGraph [more files]   ->  one JSON line per file: hash, md5, bytes"""
import hashlib
import json
import sys

import numpy as np

M = np.uint64(0x9E3779B97F4A7C15)


def file_hash(path, step=1 << 26):
    h, off, md5 = np.uint64(0), 0, hashlib.md5()
    with open(path, "rb") as f, np.errstate(over="ignore"):
        while True:
            b = f.read(step)
            if not b:
                break
            md5.update(b)
            x = np.frombuffer(b, dtype=np.uint8).astype(np.uint64)
            w = np.arange(off + 1, off + 1 + x.size, dtype=np.uint64) * M
            h = h + np.sum(x * w, dtype=np.uint64)
            off += x.size
    return {"file": path, "hash": "%016x" % int(h), "md5": md5.hexdigest(), "bytes": off}


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(json.dumps(file_hash(p)), flush=True)
