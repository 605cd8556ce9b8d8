"""Synthetic reads of the benchmark shapes, generated on the device by libkc_b200
(kc_synth_reads); see SURVEY.md 8(d).  Measurement support, not a reference interface."""
from . import _lib


def synth_reads_device(d_ptr, n_reads, read_len, genome_len=0, sub_rate=0.0, n_rate=0.0, seed=1,
                       first_read=0, zipf_loci=0, stream=None):
    """Fill device memory at d_ptr with n_reads*read_len bytes of packed reads."""
    rc = _lib.load().kc_synth_reads(d_ptr, first_read, n_reads, read_len, genome_len, sub_rate, n_rate, seed,
                                    zipf_loci, stream)
    if rc != 0:
        raise RuntimeError("kc_synth_reads failed (%d)" % rc)


# shapes of BASELINE.json's configs; genome sizes give the stated coverage
CONFIGS = {
    "c1": dict(reads=100_000, read_len=100, k=31, genome_len=1_000_000, sub_rate=0.0, n_rate=1e-3, seed=1),
    "c2": dict(reads=10_000_000, read_len=100, k=31, genome_len=100_000_000, sub_rate=1e-3, n_rate=0.0, seed=2),
    "c3": dict(reads=200_000_000, read_len=100, k=63, genome_len=1_000_000_000, sub_rate=1e-3, n_rate=0.0, seed=3),
    "c4": dict(reads=1_000_000_000, read_len=100, k=31, genome_len=3_330_000_000, sub_rate=1e-3, n_rate=0.0, seed=4),
    "c5": dict(reads=1_000_000_000, read_len=100, k=31, genome_len=3_330_000_000, sub_rate=1e-3, n_rate=0.0, seed=5,
               zipf_loci=1_000_000),
}
