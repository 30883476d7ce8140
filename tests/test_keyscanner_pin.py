"""form::KeyScanner of the host side (form_b200/host/form/keyscanner.hpp) against FORM's OWN
KeyScanner - /root/reference/form/mapping/keyscanner.{hpp,cpp} compiled unmodified into
oracle/_ref/libformref.so (it needs none of the missing libraries) - on random connection
histories: the scans handed to marginalisation must be the same, in the same order, at every
step, for the default parameters and for the corner cases of the reference's signed / unsigned
comparisons (negative limits).  CPU only."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib
import test_reference_pins as pins

_vp, _sz, _u64, _i64, _d = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int64, C.c_double
CONN = C.CFUNCTYPE(_sz, _u64, _vp)


def _bind(lib, prefix):
    for name, (res, args) in {
        "keyscanner_create": (_vp, [_i64, _i64, _sz, _d]),
        "keyscanner_destroy": (None, [_vp]),
        "keyscanner_step": (_sz, [_vp, _u64, _sz, CONN, _vp, _vp, _sz]),
        "keyscanner_size": (_sz, [_vp]),
    }.items():
        fn = getattr(lib, prefix + name)
        fn.restype, fn.argtypes = res, args
    return lib


class Scanner:
    def __init__(self, lib, prefix, params):
        self.lib, self.prefix = _bind(lib, prefix), prefix
        self.h = getattr(lib, prefix + "keyscanner_create")(*params)

    def __del__(self):
        getattr(self.lib, self.prefix + "keyscanner_destroy")(self.h)

    def step(self, idx, size, conn):
        out = np.zeros(16, dtype=np.uint64)
        n = getattr(self.lib, self.prefix + "keyscanner_step")(self.h, idx, size, CONN(lambda s, _u: conn(s)), None,
                                                              out.ctypes.data, len(out))
        return [int(v) for v in out[:n]]

    def size(self):
        return getattr(self.lib, self.prefix + "keyscanner_size")(self.h)


PARAMS = {
    "defaults": (50, 10, 10, 0.1),
    "small-window": (3, 2, 2, 0.3),
    "no-keyscan-cap": (0, 4, 5, 0.05),          # max_num_keyscans <= 0 disables the cap
    "negative-cap": (-1, 3, 4, 0.1),
    "never-age-out": (6, -1, 3, 0.1),           # size_t > int64_t(-1) is never true in the reference
    "age-out-at-once": (50, 0, 3, 0.0),
    "no-recent-scans": (5, 2, 0, 0.1),          # every scan falls out at once; ratio divides by zero
}


@pytest.mark.parametrize("name", list(PARAMS))
def test_same_marginalisation_schedule_as_forms_keyscanner(name):
    params = PARAMS[name]
    for seed in range(6):
        rng = np.random.default_rng(100 + seed)
        ours = Scanner(oracle_lib.lib(), "oracle_", params)
        theirs = Scanner(pins.ref(), "ref_", params)
        p_connected = rng.uniform(0.2, 0.9)
        for idx in range(120):
            size = int(rng.integers(1, 400))
            # one connection table per step, shared by both (the callback is asked several times
            # per scan and must answer consistently)
            table = {}

            def conn(s, table=table, rng=rng):
                if s not in table:
                    table[s] = int(rng.integers(1, 3000)) if rng.random() < p_connected else 0
                return table[s]
            # fill the table in a fixed order so both scanners see the same numbers
            for s in range(idx + 1):
                conn(s)
            a = ours.step(idx, size, conn)
            b = theirs.step(idx, size, conn)
            assert a == b, (name, seed, idx, a, b)
            assert ours.size() == theirs.size()
