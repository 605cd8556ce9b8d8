"""The C++ host surface over the C ABI: command line with the reference's flags, the
reference-shaped PrepareGPU / processKMers / FreeGPU seam, and the printer."""
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "kmer-counter_b200", "host")
CLI = os.path.join(HOST, "kmer_counter_b200")
SELFTEST = os.path.join(HOST, "shim_selftest")


def test_host_binaries_are_built():
    assert os.path.exists(CLI) and os.path.exists(SELFTEST) and os.path.exists(os.path.join(HOST, "libkc_shim.so"))


def test_printer_mode_matches_kmerprinter_format(tmp_path):
    """`print <in> <out> <k>`: every word as 32 letters, then ' count' (KMerPrinter.cpp:35-91). No GPU needed."""
    rec = np.zeros(3, dtype=[("w0", "<u8"), ("w1", "<u8"), ("cnt", "<u4")])
    rec["w0"] = [0x1b1b1b1b1b1b1b1b, 0, 2**64 - 1]
    rec["w1"] = [0xc000000000000000, 0, 1 << 62]
    rec["cnt"] = [7, 0, 4000000000]
    p = tmp_path / "r.bin"
    p.write_bytes(rec.tobytes())
    out = subprocess.run([CLI, "print", str(p), "ignored", "33"], capture_output=True, text=True, check=True).stdout
    want = "".join("%s%s %d\n" % (oracle.print_word(int(r["w0"])), oracle.print_word(int(r["w1"])), int(r["cnt"])) for r in rec)
    assert out == want


@pytest.mark.gpu
@pytest.mark.parametrize("k,method", [(31, "auto"), (31, "sort"), (63, "auto"), (28, "hash"), (96, "auto"), (100, "place")])
def test_cli_counts_a_fastq_directory(tmp_path, k, method):
    d = tmp_path / "in"
    d.mkdir()
    L, per = 100, 1500
    for i in range(3):                                        # several files; a dot-file is ignored
        (d / ("part%d.fastq" % i)).write_bytes(oracle.gen_fastq(per, L, 40000, 0.01, 0.002, seed=8, first_read=i * per))
    (d / ".hidden").write_bytes(b"@x\nAAAA\n+\nIIII\n")
    out = tmp_path / "out.bin"
    out.write_bytes(b"stale")                                 # the output is truncated, not appended to
    r = subprocess.run([CLI, "kmerLength=%d" % k, "inputFileLocation=%s" % d, "outputFile=%s" % out,
                        "gpuMemoryLimit=2000000", "noOfMergersAtOnce=3", "method=%s" % method,
                        "tempFileLocation=%s" % tmp_path, "keepRuns=1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    reads = oracle.gen_reads(3 * per, L, 40000, 0.01, 0.002, seed=8)
    assert out.read_bytes() == oracle.count(reads, L, k)
    assert "chunks=" in r.stderr and (tmp_path / "1").exists()   # several chunks went through, runs were dumped
    # each dumped run is a valid sorted run; merging them with the oracle's merger gives the artefact too
    runs = [(tmp_path / str(i)).read_bytes() for i in range(1, 50) if (tmp_path / str(i)).exists()]
    assert len(runs) >= 2 and oracle.merge_runs(runs, k) == out.read_bytes()


@pytest.mark.gpu
@pytest.mark.parametrize("k,gpus,parser", [(31, 1, "gpu"), (31, 1, "host"), (63, 1, "gpu"), (31, 2, "gpu"), (31, 4, "host"), (63, 3, "gpu")])
def test_cli_accumulates_and_counts_once(tmp_path, k, gpus, parser):
    """Default mode: chunks are only packed into super-window records, one count at the end; with
    gpus=N the chunks are dealt to N contexts and the exchanges run inside the counting kernels
    (here all contexts share device 0: KC_CLI_SAME_DEVICE). The file is the oracle's artefact."""
    d = tmp_path / "in"
    d.mkdir()
    L, per = 100, 2500
    for i in range(3):
        (d / ("part%d.fastq" % i)).write_bytes(oracle.gen_fastq(per, L, 60000, 0.01, 0.002, seed=18, first_read=i * per))
    out = tmp_path / "out.bin"
    out.write_bytes(b"stale")
    env = dict(os.environ, KC_CLI_SAME_DEVICE="1")
    r = subprocess.run([CLI, "kmerLength=%d" % k, "inputFileLocation=%s" % d, "outputFile=%s" % out,
                        "gpuMemoryLimit=1500000", "gpus=%d" % gpus, "parser=%s" % parser], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert "mode=accumulate gpus=%d parser=%s" % (gpus, parser) in r.stderr
    stats = dict(t.split("=") for t in r.stderr.split() if "=" in t and t.split("=")[1].isdigit())
    assert int(stats["chunks"]) >= 4 and int(stats["reads"]) == 3 * per
    reads = oracle.gen_reads(3 * per, L, 60000, 0.01, 0.002, seed=18)
    assert out.read_bytes() == oracle.count(reads, L, k)


@pytest.mark.gpu
def test_cli_gpu_parser_host_parser_and_fallback(tmp_path):
    L, k, R = 100, 31, 4000
    fq = oracle.gen_fastq(R, L, 50000, 0.01, 0.002, seed=12)
    want = oracle.count(oracle.gen_reads(R, L, 50000, 0.01, 0.002, seed=12), L, k)

    def run(dirname, text, *extra):
        d = tmp_path / dirname
        d.mkdir()
        (d / "reads.fastq").write_bytes(text)
        out = tmp_path / (dirname + ".bin")
        r = subprocess.run([CLI, "kmerLength=%d" % k, "inputFileLocation=%s" % d, "outputFile=%s" % out,
                            "gpuMemoryLimit=3000000", *extra], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        return out.read_bytes(), r.stderr

    got, err = run("gpu", fq)
    assert got == want and "parser=gpu" in err
    got, err = run("host", fq, "parser=host")
    assert got == want and "parser=host" in err
    got, err = run("nonl", fq[:-1])                              # no final newline
    assert got == want
    # a read of another length in the middle: the device parser refuses, the host chunker skips it
    odd = fq[:2200] + b"@odd\nACGTACGT\n+\nIIIIIIII\n" + fq[2200:]
    got, err = run("odd", odd)
    assert got == want and "parsing on the host" in err and "skipped=1" in err


@pytest.mark.gpu
@pytest.mark.parametrize("k", [31, 63, 100])
def test_cli_spills_runs_and_merges_out_of_core(tmp_path, k):
    """SURVEY N2 / C5's forced multi-run configuration: a run budget far below the distinct set, so
    merged runs leave the device for pinned host memory and the artefact is put together range by
    range (upload slices, merge-path, append)."""
    d = tmp_path / "in"
    d.mkdir()
    L, R = 100 if k < 100 else 150, 24000
    (d / "reads.fastq").write_bytes(oracle.gen_fastq(R, L, 300000, 0.01, 0.002, seed=21))
    out = tmp_path / "out.bin"
    r = subprocess.run([CLI, "kmerLength=%d" % k, "inputFileLocation=%s" % d, "outputFile=%s" % out,
                        "gpuMemoryLimit=1200000", "noOfMergersAtOnce=2", "runBudget=400000"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    stats = dict(t.split("=") for t in r.stderr.split() if "=" in t and t.split("=")[1].isdigit())
    assert int(stats["chunks"]) >= 16 and int(stats["spills"]) >= 4 and int(stats["ranges"]) >= 8, r.stderr
    want = oracle.count(oracle.gen_reads(R, L, 300000, 0.01, 0.002, seed=21), L, k)
    assert out.read_bytes() == want
    assert int(stats["records"]) * (8 * ((k + 31) // 32) + 4) == len(want)


@pytest.mark.gpu
def test_reference_shaped_seam_from_four_threads(tmp_path):
    L, k, R, chunk = 100, 31, 6000, 1000
    reads = oracle.gen_reads(R, L, 50000, 0.01, 0.002, seed=9)
    src = tmp_path / "reads.bin"
    src.write_bytes(reads.tobytes())
    out = tmp_path / "runs.bin"
    subprocess.run([SELFTEST, str(src), str(L), str(k), str(chunk), str(out)], check=True, timeout=300)
    blob = out.read_bytes()
    pos, c = 0, 0
    while pos < len(blob):
        (nb,) = struct.unpack_from("<Q", blob, pos)
        run = blob[pos + 8: pos + 8 + nb]
        assert run == oracle.process_chunk(reads[c * chunk * L:(c + 1) * chunk * L], L, k), c
        pos += 8 + nb
        c += 1
    assert c == R // chunk
