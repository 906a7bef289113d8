"""
CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this package.  ``mdhelper_b200`` never
does: the product path has no CPU fallback.

``oracle.lib()`` returns the ctypes handle of ``libmdh_oracle.so`` (the C
restatement in ``mdh_oracle.c``), building it with ``make`` on first use.
"""

import ctypes
import pathlib
import subprocess

_HERE = pathlib.Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> pathlib.Path:
    so = _HERE / "libmdh_oracle.so"
    src = _HERE / "mdh_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s", "-B"], check=True)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(str(build()))
        i64, f64 = ctypes.c_int64, ctypes.c_double
        p = ctypes.c_void_p
        L.mdho_capped_distance_bruteforce.restype = i64
        L.mdho_capped_distance_bruteforce.argtypes = [
            p, i64, i64, p, i64, p, f64, f64, p, p, i64]
        L.mdho_capped_distance_cells.restype = i64
        L.mdho_capped_distance_cells.argtypes = [
            p, i64, p, i64, p, f64, f64, p, p, i64]
        L.mdho_ortho_pbc.restype = None
        L.mdho_ortho_pbc.argtypes = [p, i64, p]
        L.mdho_capped_distance_triclinic.restype = i64
        L.mdho_capped_distance_triclinic.argtypes = [
            p, i64, p, i64, p, f64, f64, p, p, i64]
        L.mdho_delta_fourier_transform_sum.restype = None
        L.mdho_delta_fourier_transform_sum.argtypes = [
            p, i64, p, i64, p, ctypes.c_int]
        L.mdho_max_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB
