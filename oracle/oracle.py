"""ctypes loader for the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  PARITY UNPINNED at the limb level -- see oracle/fheram_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent


class Params(C.Structure):
    _fields_ = [
        ("log_n", C.c_int32), ("base2k", C.c_int32), ("k_pt", C.c_int32), ("k_ct", C.c_int32),
        ("k_addr", C.c_int32), ("k_evk_trace", C.c_int32), ("k_evk_ggsw_inv", C.c_int32),
        ("word_size", C.c_int32), ("n_decomp", C.c_int32), ("decomp_n", C.c_int32 * 8),
        ("max_addr", C.c_uint64),
    ]


def build(force: bool = False) -> None:
    """Compile both oracle variants (gcc, see oracle/Makefile)."""
    need = force or any(not (_DIR / f).exists() or
                        (_DIR / f).stat().st_mtime < (_DIR / "fheram_oracle.c").stat().st_mtime
                        for f in ("liboracle_exact.so", "liboracle_fft64.so"))
    if need:
        subprocess.check_call(["make", "-C", str(_DIR), "-s"] + (["-B"] if force else []))


_P64 = C.POINTER(C.c_int64)
_PU8 = C.POINTER(C.c_uint8)


def _p(a: np.ndarray):
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_P64)


class Oracle:
    """Thin object wrapper over liboracle_{exact,fft64}.so."""

    def __init__(self, params: Params | None = None, backend: str = "exact", **overrides):
        so = _DIR / f"liboracle_{backend}.so"
        if not so.exists():
            build()
        self.lib = lib = C.CDLL(str(so))
        self.backend = backend
        V = C.c_void_p
        sig = {
            "orc_params_snapshot": (None, [C.POINTER(Params)]),
            "orc_params_readme": (None, [C.POINTER(Params)]),
            "orc_get_base_2d": (C.c_int, [C.c_uint32, C.POINTER(C.c_int32), C.c_int,
                                          C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
            "orc_base1d_max": (C.c_uint32, [C.POINTER(C.c_int32), C.c_int]),
            "orc_base1d_gap": (C.c_uint32, [C.POINTER(C.c_int32), C.c_int, C.c_int]),
            "orc_base1d_decomp": (None, [C.POINTER(C.c_int32), C.c_int, C.c_uint32, _PU8]),
            "orc_base1d_recomp": (C.c_uint32, [C.POINTER(C.c_int32), C.c_int, _PU8]),
            "orc_reverse_bits_msb": (C.c_uint64, [C.c_uint64, C.c_uint32]),
            "orc_cast_u8_to_signed": (C.c_int64, [C.c_uint8, C.c_int]),
            "orc_ctx_new": (V, [C.POINTER(Params)]),
            "orc_ctx_free": (None, [V]),
            "orc_backend_name": (C.c_char_p, []),
            "orc_n": (C.c_size_t, [V]), "orc_glwe_len": (C.c_size_t, [V]),
            "orc_ggsw_len": (C.c_size_t, [V]), "orc_atk_len": (C.c_size_t, [V]),
            "orc_evk_inv_len": (C.c_size_t, [V]), "orc_n_gal": (C.c_int, [V]),
            "orc_n_ggsw": (C.c_int, [V]), "orc_n_glwe_per_subram": (C.c_int, [V]),
            "orc_gal_el": (C.c_int64, [V, C.c_int]),
            "orc_source_new": (V, [_PU8]), "orc_source_free": (None, [V]),
            "orc_source_next_u64": (C.c_uint64, [V]), "orc_source_next_u32": (C.c_uint32, [V]),
            "orc_source_fill_bytes": (None, [V, _PU8, C.c_size_t]),
            "orc_secret_gen": (None, [V, V, _P64]),
            "orc_keygen": (None, [V, _P64, V, V, _P64, _P64, _P64]),
            "orc_ram_encrypt": (None, [V, _PU8, _P64, V, V, _P64]),
            "orc_address_encrypt": (None, [V, C.c_uint32, _P64, V, V, _P64]),
            "orc_encrypt_byte": (None, [V, C.c_uint8, _P64, V, V, _P64]),
            "orc_decrypt_glwe": (None, [V, _P64, _P64, C.c_int64, _P64, C.POINTER(C.c_double)]),
            "orc_glwe_decrypt": (None, [V, _P64, _P64, _P64]),
            "orc_ggsw_decrypt_row": (None, [V, _P64, C.c_int, C.c_int, _P64, _P64]),
            "orc_keys_prepare": (V, [V, _P64, _P64, _P64]),
            "orc_keys_free": (None, [V]),
            "orc_glwe_normalize": (None, [V, _P64]),
            "orc_glwe_rsh": (None, [V, C.c_int, _P64]),
            "orc_glwe_rotate": (None, [V, C.c_int64, _P64, _P64]),
            "orc_glwe_small_automorphism": (None, [V, C.c_int64, _P64, _P64]),
            "orc_external_product": (None, [V, _P64, _P64, _P64]),
            "orc_coordinate_product": (None, [V, _P64, _P64, C.c_int, _P64]),
            "orc_automorphism": (None, [V, V, C.c_int, C.c_int, _P64, _P64]),
            "orc_trace": (None, [V, V, C.c_int, C.c_int, _P64, _P64]),
            "orc_packer_new": (V, [V]), "orc_packer_free": (None, [V]),
            "orc_packer_add": (None, [V, V, _P64]), "orc_packer_flush": (None, [V, _P64]),
            "orc_ggsw_automorphism_inv": (None, [V, V, _P64, _P64]),
            "orc_packer_combine": (None, [V, V, C.c_int, _P64, _P64]),
            "orc_ram_new": (V, [V]), "orc_ram_free": (None, [V]),
            "orc_ram_load": (None, [V, _P64]), "orc_ram_store": (None, [V, _P64]),
            "orc_ram_tree_store": (None, [V, _P64]), "orc_ram_state": (C.c_int, [V]),
            "orc_set_ram_threads": (None, [C.c_int]),
            "orc_external_product_many": (None, [V, _P64, C.c_int, _P64, _P64, C.c_int]),
            "orc_ram_read": (C.c_int, [V, _P64, V, _P64]),
            "orc_ram_read_prepare_write": (C.c_int, [V, _P64, V, _P64]),
            "orc_ram_write": (C.c_int, [V, _P64, _P64, V]),
            "orc_ram_read_many": (C.c_int, [V, _P64, C.c_int, V, _P64, C.c_int]),
            "orc_op_counters": (None, [V, C.POINTER(C.c_uint64)]),
        }
        for name, (res, args) in sig.items():
            f = getattr(lib, name)
            f.restype, f.argtypes = res, args
        if params is None:
            params = Params()
            lib.orc_params_snapshot(C.byref(params))
        for k, v in overrides.items():
            if k == "decomp_n":
                params.n_decomp = len(v)
                for i, d in enumerate(v):
                    params.decomp_n[i] = d
            else:
                setattr(params, k, v)
        self.params = params
        self.ctx = lib.orc_ctx_new(C.byref(params))
        self.n = lib.orc_n(self.ctx)
        self.glwe_len = lib.orc_glwe_len(self.ctx)
        self.ggsw_len = lib.orc_ggsw_len(self.ctx)
        self.atk_len = lib.orc_atk_len(self.ctx)
        self.evk_inv_len = lib.orc_evk_inv_len(self.ctx)
        self.n_gal = lib.orc_n_gal(self.ctx)
        self.n_ggsw = lib.orc_n_ggsw(self.ctx)
        self.n_glwe = lib.orc_n_glwe_per_subram(self.ctx)
        self.word_size = params.word_size
        self.size_ct = -(-params.k_ct // params.base2k)

    # ---- params helpers -------------------------------------------------------------
    @staticmethod
    def readme_params() -> Params:
        lib = C.CDLL(str(_DIR / "liboracle_exact.so"))
        p = Params()
        lib.orc_params_readme(C.byref(p))
        return p

    def gal_els(self):
        return [self.lib.orc_gal_el(self.ctx, i) for i in range(self.n_gal)]

    def base2d(self, max_addr=None, decomp=None):
        decomp = list(self.params.decomp_n[: self.params.n_decomp]) if decomp is None else decomp
        max_addr = self.params.max_addr if max_addr is None else max_addr
        arr = (C.c_int32 * len(decomp))(*decomp)
        lens = (C.c_int32 * 8)()
        digits = (C.c_int32 * 64)()
        n = self.lib.orc_get_base_2d(max_addr, arr, len(decomp), lens, digits)
        return [[digits[i * 8 + j] for j in range(lens[i])] for i in range(n)]

    # ---- sources ---------------------------------------------------------------------
    def source(self, seed):
        if isinstance(seed, int):
            seed = bytes([seed] * 32)
        buf = (C.c_uint8 * 32)(*seed)
        return self.lib.orc_source_new(buf)

    def source_bytes(self, src, n):
        out = np.zeros(n, dtype=np.uint8)
        self.lib.orc_source_fill_bytes(src, out.ctypes.data_as(_PU8), n)
        return out

    def source_u32(self, src):
        return self.lib.orc_source_next_u32(src)

    # ---- client ----------------------------------------------------------------------
    def secret_gen(self, xs):
        sk = np.zeros(self.n, dtype=np.int64)
        self.lib.orc_secret_gen(self.ctx, xs, _p(sk))
        return sk

    def keygen(self, sk, xa, xe):
        atk = np.zeros(self.n_gal * self.atk_len, dtype=np.int64)
        tsk = np.zeros(self.evk_inv_len, dtype=np.int64)
        inv = np.zeros(self.evk_inv_len, dtype=np.int64)
        self.lib.orc_keygen(self.ctx, _p(sk), xa, xe, _p(atk), _p(tsk), _p(inv))
        return atk, tsk, inv

    def ram_encrypt(self, data, sk, xa, xe):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        assert data.size == self.params.max_addr * self.word_size
        out = np.zeros(self.word_size * self.n_glwe * self.glwe_len, dtype=np.int64)
        self.lib.orc_ram_encrypt(self.ctx, data.ctypes.data_as(_PU8), _p(sk), xa, xe, _p(out))
        return out

    def address_encrypt(self, value, sk, xa, xe):
        out = np.zeros(self.n_ggsw * self.ggsw_len, dtype=np.int64)
        self.lib.orc_address_encrypt(self.ctx, int(value), _p(sk), xa, xe, _p(out))
        return out

    def encrypt_byte(self, value, sk, xa, xe):
        out = np.zeros(self.glwe_len, dtype=np.int64)
        self.lib.orc_encrypt_byte(self.ctx, int(value), _p(sk), xa, xe, _p(out))
        return out

    def decrypt_glwe(self, glwe, sk, want):
        v = C.c_int64()
        noise = C.c_double()
        glwe = np.ascontiguousarray(glwe, dtype=np.int64)
        self.lib.orc_decrypt_glwe(self.ctx, _p(glwe), _p(sk), int(want), C.byref(v), C.byref(noise))
        return v.value, noise.value

    def glwe_decrypt(self, glwe, sk):
        pt = np.zeros(self.size_ct * self.n, dtype=np.int64)
        glwe = np.ascontiguousarray(glwe, dtype=np.int64)
        self.lib.orc_glwe_decrypt(self.ctx, _p(glwe), _p(sk), _p(pt))
        return pt.reshape(self.size_ct, self.n)

    def cast_u8_to_signed(self, v, bits):
        return self.lib.orc_cast_u8_to_signed(int(v), bits)

    # ---- keys / ops ------------------------------------------------------------------
    def keys_prepare(self, atk, tsk, inv):
        return self.lib.orc_keys_prepare(self.ctx, _p(atk), _p(tsk), _p(inv))

    def external_product(self, glwe, ggsw):
        out = np.zeros(self.glwe_len, dtype=np.int64)
        self.lib.orc_external_product(self.ctx, _p(np.ascontiguousarray(glwe)), _p(np.ascontiguousarray(ggsw)), _p(out))
        return out

    def external_product_many(self, glwes, ggsw, threads=1):
        glwes = np.ascontiguousarray(glwes, dtype=np.int64)
        n = glwes.size // self.glwe_len
        out = np.zeros((n, self.glwe_len), dtype=np.int64)
        self.lib.orc_external_product_many(self.ctx, _p(glwes.reshape(-1)), n, _p(np.ascontiguousarray(ggsw)), _p(out.reshape(-1)), threads)
        return out

    def coordinate_product(self, glwe, ggsws, n):
        out = np.zeros(self.glwe_len, dtype=np.int64)
        self.lib.orc_coordinate_product(self.ctx, _p(np.ascontiguousarray(glwe)), _p(np.ascontiguousarray(ggsws)), n, _p(out))
        return out

    def automorphism(self, keys, gal_idx, mode, glwe):
        out = np.zeros(self.glwe_len, dtype=np.int64)
        self.lib.orc_automorphism(self.ctx, keys, gal_idx, mode, _p(np.ascontiguousarray(glwe)), _p(out))
        return out

    def trace(self, keys, glwe, start=0, end=None):
        out = np.zeros(self.glwe_len, dtype=np.int64)
        end = self.n_gal if end is None else end
        self.lib.orc_trace(self.ctx, keys, start, end, _p(np.ascontiguousarray(glwe)), _p(out))
        return out

    def glwe_rsh(self, k, glwe):
        out = np.array(glwe, dtype=np.int64, copy=True)
        self.lib.orc_glwe_rsh(self.ctx, k, _p(out))
        return out

    def glwe_normalize(self, glwe):
        out = np.array(glwe, dtype=np.int64, copy=True)
        self.lib.orc_glwe_normalize(self.ctx, _p(out))
        return out

    def glwe_rotate(self, k, glwe):
        out = np.zeros(self.glwe_len, dtype=np.int64)
        self.lib.orc_glwe_rotate(self.ctx, k, _p(np.ascontiguousarray(glwe)), _p(out))
        return out

    def glwe_small_automorphism(self, p, glwe):
        out = np.zeros(self.glwe_len, dtype=np.int64)
        self.lib.orc_glwe_small_automorphism(self.ctx, p, _p(np.ascontiguousarray(glwe)), _p(out))
        return out

    def pack(self, keys, inputs):
        """inputs: list of N entries, each a GLWE array or None (GLWEPacker add x N, flush)."""
        pk = self.lib.orc_packer_new(self.ctx)
        for g in inputs:
            self.lib.orc_packer_add(pk, keys, None if g is None else _p(np.ascontiguousarray(g)))
        out = np.zeros(self.glwe_len, dtype=np.int64)
        self.lib.orc_packer_flush(pk, _p(out))
        self.lib.orc_packer_free(pk)
        return out

    def packer_combine(self, keys, level, a, b=None):
        """GLWEPacker::combine at `level`; returns the updated accumulator."""
        out = np.array(a, dtype=np.int64, copy=True)
        self.lib.orc_packer_combine(self.ctx, keys, level, _p(out), None if b is None else _p(np.ascontiguousarray(b)))
        return out

    def ggsw_automorphism_inv(self, keys, ggsw):
        out = np.zeros(self.ggsw_len, dtype=np.int64)
        self.lib.orc_ggsw_automorphism_inv(self.ctx, keys, _p(np.ascontiguousarray(ggsw)), _p(out))
        return out

    def ggsw_decrypt_row(self, ggsw, row, col_in, sk):
        size_addr = -(-self.params.k_addr // self.params.base2k)
        pt = np.zeros(size_addr * self.n, dtype=np.int64)
        self.lib.orc_ggsw_decrypt_row(self.ctx, _p(np.ascontiguousarray(ggsw)), row, col_in, _p(sk), _p(pt))
        return pt.reshape(size_addr, self.n)

    # ---- ram -------------------------------------------------------------------------
    def ram_new(self, cts=None):
        r = self.lib.orc_ram_new(self.ctx)
        if cts is not None:
            self.lib.orc_ram_load(r, _p(cts))
        return r

    def ram_store(self, ram):
        out = np.zeros(self.word_size * self.n_glwe * self.glwe_len, dtype=np.int64)
        self.lib.orc_ram_store(ram, _p(out))
        return out

    def ram_tree_store(self, ram):
        out = np.zeros(self.word_size * self.glwe_len, dtype=np.int64)
        self.lib.orc_ram_tree_store(ram, _p(out))
        return out

    def set_ram_threads(self, n: int):
        """sub-RAMs of one read / read_prepare_write / write on n threads (same results; checker speed only)"""
        self.lib.orc_set_ram_threads(int(n))

    def ram_read(self, ram, addr, keys):
        out = np.zeros(self.word_size * self.glwe_len, dtype=np.int64)
        rc = self.lib.orc_ram_read(ram, _p(addr), keys, _p(out))
        return rc, out.reshape(self.word_size, self.glwe_len)

    def ram_read_prepare_write(self, ram, addr, keys):
        out = np.zeros(self.word_size * self.glwe_len, dtype=np.int64)
        rc = self.lib.orc_ram_read_prepare_write(ram, _p(addr), keys, _p(out))
        return rc, out.reshape(self.word_size, self.glwe_len)

    def ram_write(self, ram, w, addr, keys):
        w = np.ascontiguousarray(w, dtype=np.int64)
        return self.lib.orc_ram_write(ram, _p(w), _p(addr), keys)

    def ram_read_many(self, ram, addrs, n, keys, threads):
        out = np.zeros(n * self.word_size * self.glwe_len, dtype=np.int64)
        rc = self.lib.orc_ram_read_many(ram, _p(addrs), n, keys, _p(out), threads)
        return rc, out.reshape(n, self.word_size, self.glwe_len)

    def op_counters(self):
        out = (C.c_uint64 * 2)()
        self.lib.orc_op_counters(self.ctx, out)
        return out[0], out[1]


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
