"""A/B timing of icl_host_read_m8 between builds of libicl_host.so (alternating runs, best of N: the build container has
noisy neighbours and single runs differ by 30 %).

    python tools/host_ab.py <dir with blast.m8 + genome1.gff> <threads> <repeats> <lib A.so> <lib B.so> ...
"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wgdi_b200 import _host, base                                 # noqa: E402


def main():
    d, threads, reps, libs = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4:]
    b1, o1 = _host.pack_strings(base.newgff(os.path.join(d, "genome1.gff")).index)
    p1 = C.cast(C.c_char_p(b1), C.c_void_p)
    loaded = {}
    for path in libs:
        lib = C.CDLL(path)
        lib.icl_host_read_m8.restype = C.c_int
        lib.icl_host_read_m8.argtypes = [C.c_char_p, C.c_int32, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_char_p, C.c_int32]
        loaded[path] = lib
    best, kept = {p: float("inf") for p in libs}, {}
    for _ in range(reps):
        for path, lib in loaded.items():
            out, err = _host.icl_hits(), C.create_string_buffer(512)
            t = time.perf_counter()
            rc = lib.icl_host_read_m8(os.path.join(d, "blast.m8").encode(), 0, 100.0, 1e-5, p1, o1.ctypes.data, len(o1) - 1,
                                      p1, o1.ctypes.data, len(o1) - 1, threads, C.byref(out), err, 512)
            best[path] = min(best[path], time.perf_counter() - t)
            assert rc == 0, err.value
            kept[path] = int(out.n)
            lib.icl_host_free_hits(C.byref(out))
    for p in libs:
        print(f"{p}: best {best[p] * 1e3:.1f} ms, {kept[p]} rows kept")


if __name__ == "__main__":
    main()
