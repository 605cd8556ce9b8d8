"""ctypes bindings for the CPU oracle (oracle/kc_oracle.c) and, when built, for
oracle/_ref (the reference's own sources compiled in place).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's CPU legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libkc_oracle.so")
_REF_LIB = os.path.join(_HERE, "_ref", "libkc_ref.so")
REF_BIN = os.path.join(_HERE, "_ref", "ref_count")

u8p = C.POINTER(C.c_ubyte)


def build(ref=True):
    """Compile the C restatement and (where /root/reference exists) oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", _HERE])
    if ref:
        subprocess.check_call(["bash", os.path.join(_HERE, "build_ref.sh")])


def _load():
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "kc_oracle.c")):
        build(ref=False)
    lib = C.CDLL(_LIB)
    lib.kco_words.restype = C.c_uint32
    lib.kco_record_size.restype = C.c_uint32
    lib.kco_output_size.restype = C.c_uint64
    lib.kco_output_size.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
    lib.kco_encode_read.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.kco_extract_read.restype = C.c_uint32
    lib.kco_extract_read.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.kco_process_chunk.restype = C.c_int64
    lib.kco_process_chunk.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
    lib.kco_merge_runs.restype = C.c_uint64
    lib.kco_merge_runs.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_uint32, C.c_uint32, C.c_void_p]
    for name in ("kco_count",):
        f = getattr(lib, name)
        f.restype = C.c_int64
        f.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_void_p)]
    lib.kco_count_mt.restype = C.c_int64
    lib.kco_count_mt.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32,
                                 C.POINTER(C.c_void_p)]
    lib.kco_naive_count.restype = C.c_int64
    lib.kco_naive_count.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_void_p)]
    lib.kco_free.argtypes = [C.c_void_p]
    lib.kco_parse_fastq.restype = C.c_int64
    lib.kco_parse_fastq.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.kco_print_word.argtypes = [C.c_uint64, C.c_char_p]
    lib.kco_splitmix64.restype = C.c_uint64
    lib.kco_splitmix64.argtypes = [C.c_uint64]
    lib.kco_gen_reads.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.c_double,
                                  C.c_double, C.c_uint64]
    lib.kco_gen_reads_zipf.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.c_double,
                                       C.c_double, C.c_uint64, C.c_uint64, C.c_double]
    lib.kco_fastq_bytes.restype = C.c_uint64
    lib.kco_fastq_bytes.argtypes = [C.c_uint64, C.c_uint32]
    lib.kco_gen_fastq.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.c_double,
                                  C.c_double, C.c_uint64]
    return lib


lib = _load()


def _as_u8(reads):
    if isinstance(reads, (bytes, bytearray)):
        return np.frombuffer(bytes(reads), dtype=np.uint8)
    return np.ascontiguousarray(reads, dtype=np.uint8)


def words(k):
    return (k + 31) // 32


def record_size(k):
    return 8 * words(k) + 4


def encode_read(read: bytes):
    L = len(read)
    w = np.zeros((L + 31) // 32 + 1, dtype=np.uint64)
    f = np.zeros((L + 63) // 64, dtype=np.uint64)
    rc = lib.kco_encode_read(read, L, w.ctypes.data, f.ctypes.data)
    if rc != 0:
        raise ValueError("read length %d is not representable by the reference" % L)
    return w[: (L + 31) // 32], f


def extract_read(read: bytes, k):
    """Raw per-read slots as the reference kernel leaves them: (L-k+1) records."""
    L = len(read)
    w = np.zeros((L + 31) // 32 + 1, dtype=np.uint64)
    f = np.zeros((L + 63) // 64, dtype=np.uint64)
    if lib.kco_encode_read(read, L, w.ctypes.data, f.ctypes.data) != 0:
        raise ValueError("bad L")
    out = np.zeros((L - k + 1) * record_size(k), dtype=np.uint8)
    n = lib.kco_extract_read(w.ctypes.data, f.ctypes.data, L, k, out.ctypes.data)
    return out, n


def process_chunk(reads, L, k, do_sort=True) -> bytes:
    a = _as_u8(reads)
    cap = lib.kco_output_size(a.size, L, k)
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    n = lib.kco_process_chunk(a.ctypes.data, a.size, L, k, 1 if do_sort else 0, out.ctypes.data)
    if n < 0:
        raise ValueError("oracle rejected (L=%d, k=%d)" % (L, k))
    return out[:n].tobytes()


def merge_runs(runs, k) -> bytes:
    arrs = [_as_u8(r) for r in runs]
    n = len(arrs)
    ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrs])
    lens = (C.c_uint64 * max(n, 1))(*[a.size for a in arrs])
    out = np.zeros(max(sum(a.size for a in arrs), 1), dtype=np.uint8)
    nb = lib.kco_merge_runs(ptrs, lens, n, k, out.ctypes.data)
    return out[:nb].tobytes()


def _take(ptr, n):
    if n < 0:
        raise ValueError("oracle rejected the input")
    data = C.string_at(ptr, n) if n else b""
    lib.kco_free(ptr)
    return data


def count(reads, L, k, chunk_reads=0, threads=1) -> bytes:
    """The parity artefact: final merged sorted-unique record file, as bytes."""
    a = _as_u8(reads)
    p = C.c_void_p()
    n = lib.kco_count_mt(a.ctypes.data, a.size, L, k, chunk_reads, threads, C.byref(p))
    return _take(p, n)


def naive_count(reads, L, k, strict=False) -> bytes:
    a = _as_u8(reads)
    p = C.c_void_p()
    n = lib.kco_naive_count(a.ctypes.data, a.size, L, k, 1 if strict else 0, C.byref(p))
    return _take(p, n)


def parse_fastq(text: bytes, cap=None) -> bytes:
    a = _as_u8(text)
    cap = cap if cap is not None else a.size + 1
    out = np.zeros(cap + 1, dtype=np.uint8)
    n = lib.kco_parse_fastq(a.ctypes.data, a.size, out.ctypes.data, cap)
    return out[:n].tobytes()


def print_word(w) -> str:
    buf = C.create_string_buffer(32)
    lib.kco_print_word(int(w), buf)
    return buf.raw.decode()


def gen_reads(n_reads, L, genome_len=0, sub_rate=0.0, n_rate=0.0, seed=1, first_read=0,
              zipf_loci=0, zipf_s=1.2) -> np.ndarray:
    out = np.empty(n_reads * L, dtype=np.uint8)
    lib.kco_gen_reads_zipf(out.ctypes.data, first_read, n_reads, L, genome_len, sub_rate, n_rate, seed,
                           zipf_loci, zipf_s)
    return out


def gen_fastq(n_reads, L, genome_len=0, sub_rate=0.0, n_rate=0.0, seed=1, first_read=0) -> bytes:
    out = np.empty(lib.kco_fastq_bytes(n_reads, L), dtype=np.uint8)
    lib.kco_gen_fastq(out.ctypes.data, first_read, n_reads, L, genome_len, sub_rate, n_rate, seed)
    return out.tobytes()


def records_to_arrays(data: bytes, k):
    """Packed records -> (keys[n, W] uint64, counts[n] uint32)."""
    W, S = words(k), record_size(k)
    a = np.frombuffer(data, dtype=np.uint8).reshape(-1, S)
    keys = np.ascontiguousarray(a[:, : 8 * W]).view(np.uint64).reshape(-1, W)
    counts = np.ascontiguousarray(a[:, 8 * W:]).view(np.uint32).reshape(-1)
    return keys, counts


# ----------------------------------------------------------------- oracle/_ref

_ref = None


def ref_available():
    return os.path.exists(_REF_LIB)


def ref():
    """The reference's own code (oracle/_ref/libkc_ref.so); None when not built."""
    global _ref
    if _ref is None and ref_available():
        r = C.CDLL(_REF_LIB)
        r.ref_process_chunk.restype = C.c_int64
        r.ref_process_chunk.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_void_p]
        r.ref_output_size.restype = C.c_uint64
        r.ref_output_size.argtypes = [C.c_uint64, C.c_int64, C.c_int64]
        r.ref_dump_run.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_int64]
        r.ref_merge_files.argtypes = [C.POINTER(C.c_char_p), C.c_uint32, C.c_char_p, C.c_uint64]
        r.ref_count_packed.restype = C.c_int64
        r.ref_count_packed.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_uint64, C.c_uint32,
                                       C.c_char_p, C.c_char_p]
        if hasattr(r, "ref_count_packed_ex"):
            r.ref_count_packed_ex.restype = C.c_int64
            r.ref_count_packed_ex.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_uint64, C.c_uint32,
                                              C.c_uint32, C.c_uint32, C.c_char_p, C.c_char_p, C.POINTER(C.c_double)]
        r.ref_read_fastq_dir.restype = C.c_int64
        r.ref_read_fastq_dir.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        r.ref_free.argtypes = [C.c_void_p]
        _ref = r
    return _ref


def ref_process_chunk(reads, L, k, do_sort=True) -> bytes:
    a = _as_u8(reads)
    r = ref()
    cap = r.ref_output_size(a.size, L, k)
    out = np.zeros(max(cap, 1) + 64, dtype=np.uint8)
    n = r.ref_process_chunk(a.ctypes.data, a.size, L, k, 1 if do_sort else 0, out.ctypes.data)
    return out[:n].tobytes()


def ref_merge_files(paths, out_path, k):
    arr = (C.c_char_p * max(len(paths), 1))(*[p.encode() for p in paths])
    ref().ref_merge_files(arr, len(paths), out_path.encode(), k)


def ref_count_packed(reads, L, k, chunk_reads, threads, tmp_dir, out_path) -> int:
    a = _as_u8(reads)
    return ref().ref_count_packed(a.ctypes.data, a.size, L, k, chunk_reads, threads, tmp_dir.encode(),
                                  out_path.encode())


def ref_count_packed_ex(reads, L, k, chunk_reads, threads, fan_in, merge_threads, tmp_dir, out_path) -> dict:
    """Pipeline B with per-stage seconds and the reference's merge policy (fan-in, merger threads)."""
    a = _as_u8(reads)
    t = (C.c_double * 6)()
    runs = ref().ref_count_packed_ex(a.ctypes.data, a.size, L, k, chunk_reads, threads, fan_in, merge_threads,
                                     tmp_dir.encode(), out_path.encode(), t)
    return {"runs": int(runs), "encode_extract_thread_s": t[0], "sort_thread_s": t[1], "reduce_dump_thread_s": t[2],
            "chunk_phase_wall_s": t[3], "merge_wall_s": t[4], "threads_busy": int(t[5])}


def ref_read_fastq_dir(path, chunk_size=1 << 30):
    p = C.c_void_p()
    L = C.c_int64()
    n = ref().ref_read_fastq_dir(path.encode(), chunk_size, C.byref(p), C.byref(L))
    data = C.string_at(p, n) if n else b""
    ref().ref_free(p)
    return data, L.value
