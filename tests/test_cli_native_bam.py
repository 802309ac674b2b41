"""The CLI shell on a REAL .bam (written by tests/bam_writer.py) through the native decoder:
VCF body identical to what the reference's own main_ctrl wrote for the same records
(tests/golden/cli_dataset1.json / cli_config1.json, made by oracle/gen_cli_golden.py).
CPU variant: kernels replaced by the pipeline emulator; -m gpu variant: the CUDA path."""
import json
import os
import pickle

import pytest

import bam_writer
import golden_util
from cutesv_b200 import bamio, cli
from oracle import gen_cli_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _to_real_bam(pickled, path):
    ds = pickle.load(open(pickled, "rb"))
    order = {n: i for i, (n, _) in enumerate(ds["contigs"])}
    reads = sorted(ds["reads"], key=lambda r: (order[r.reference_name], r.reference_start))  # stable, like an indexed BAM
    bam_writer.write_bam(path, ds["contigs"], reads, extra_unmapped=2)
    return path


def _run(engine, tmp_path, which, extra=()):
    bamio.build()
    if which in (1, 3):
        gold = json.load(open(os.path.join(golden_util.GOLDEN, "cli_dataset1.json" if which == 1 else "cli_dataset1_bed.json")))
        pk, fa, out, wd = gen_cli_golden.materialise(str(tmp_path))
        if which == 3:   # -include_bed: the read filter of cuteSV:717-723 with window-assigned regions
            extra = list(extra) + ["-include_bed", gen_cli_golden.write_bed(str(tmp_path))]
    else:
        gold = json.load(open(os.path.join(golden_util.GOLDEN, "cli_config1.json")))
        pk, fa, out, wd = gen_cli_golden.materialise_config1(str(tmp_path))
    bam = _to_real_bam(pk, str(tmp_path / "real.bam"))
    assert bamio.is_bam(bam)
    argv = [bam, fa, out, wd] + gold["flags"] + list(extra)
    cli.main_ctrl(cli.build_parser().parse_args(argv), argv, engine=engine)
    return [l for l in open(out) if not l.startswith("##")], gold["lines"], wd


def test_cli_native_bam_dataset1_cpu(tmp_path):
    from emul_engine import EmulEngine
    lines, gold, wd = _run(EmulEngine(), tmp_path, 1, ["--retain_work_dir"])
    assert lines == gold
    assert os.path.exists(os.path.join(wd, "reads.pickle"))
    st = cli.main_ctrl.last_stages   # the stage split of the wall time that scripts/bench_cli.py reports
    assert {"scan", "names_and_ties", "cluster_and_fetch", "rows", "vcf", "scan.csv_extract_append", "scan.ins_sequences"} <= set(st)
    assert all(v >= 0 for v in st.values()) and st["scan"] >= st["scan.csv_extract_append"]


def test_cli_native_bam_config1_cpu(tmp_path):
    from emul_engine import EmulEngine
    lines, gold, _ = _run(EmulEngine(), tmp_path, 2)
    assert len(lines) > 250 and lines == gold


@pytest.mark.parametrize("name", ["cli_dataset1_hifi_readid", "cli_dataset1_nogt_noseq"])
def test_cli_native_bam_flag_sets_cpu(tmp_path, name):
    """Other flag sets (RNAMES, no genotype / no sequences, size limits, merge thresholds, max_split_parts):
    VCF body identical to the reference's own run with the same flags."""
    from emul_engine import EmulEngine
    bamio.build()
    gold = json.load(open(os.path.join(golden_util.GOLDEN, name + ".json")))
    pk, fa, out, wd = gen_cli_golden.materialise(str(tmp_path))
    bam = _to_real_bam(pk, str(tmp_path / "real.bam"))
    argv = [bam, fa, out, wd] + gold["flags"]
    cli.main_ctrl(cli.build_parser().parse_args(argv), argv, engine=EmulEngine())
    import vcf_util
    assert vcf_util.normalise_rnames([l for l in open(out) if not l.startswith("##")]) == vcf_util.normalise_rnames(gold["lines"])


def test_cli_native_bam_include_bed_cpu(tmp_path):
    from emul_engine import EmulEngine
    lines, gold, _ = _run(EmulEngine(), tmp_path, 3)
    assert 5 < len(lines) < 28 and lines == gold


def test_cli_pysam_path_include_bed_cpu(tmp_path, monkeypatch):
    """Same golden through the pysam-shaped source (window-by-window fetch) with the test-only fake pysam."""
    import sys
    from emul_engine import EmulEngine
    monkeypatch.syspath_prepend(os.path.join(ROOT, "tests", "fake_pysam"))
    sys.modules.pop("pysam", None)
    gold = json.load(open(os.path.join(golden_util.GOLDEN, "cli_dataset1_bed.json")))
    bam, fa, out, wd = gen_cli_golden.materialise(str(tmp_path))
    argv = [bam, fa, out, wd] + gold["flags"] + ["-include_bed", gen_cli_golden.write_bed(str(tmp_path))]
    cli.main_ctrl(cli.build_parser().parse_args(argv), argv, engine=EmulEngine())
    assert [l for l in open(out) if not l.startswith("##")] == gold["lines"]
    sys.modules.pop("pysam", None)


@pytest.mark.gpu
@pytest.mark.parametrize("which", [1, 2, 3])
def test_cli_native_bam_gpu(engine, tmp_path, which):
    lines, gold, _ = _run(engine, tmp_path, which)
    assert lines == gold


@pytest.mark.parametrize("which,packet", [(2, 700), (1, 97)])
def test_cli_native_bam_many_packets_cpu(tmp_path, monkeypatch, which, packet):
    """Small extraction packets (several csv_extract calls, provisional read ids and alignment chunks spanning packets):
    the VCF body must not depend on the packet size."""
    from emul_engine import EmulEngine
    monkeypatch.setattr(cli, "PACKET_READS", packet)
    lines, gold, _ = _run(EmulEngine(), tmp_path, which)
    assert lines == gold


def test_cli_pysam_path_many_packets_cpu(tmp_path, monkeypatch):
    """The pysam-shaped source (window-by-window fetch through the test-only fake pysam) with small packets."""
    import sys
    from emul_engine import EmulEngine
    monkeypatch.syspath_prepend(os.path.join(ROOT, "tests", "fake_pysam"))
    monkeypatch.setattr(cli, "PACKET_READS", 211)
    sys.modules.pop("pysam", None)
    gold = json.load(open(os.path.join(golden_util.GOLDEN, "cli_dataset1.json")))
    bam, fa, out, wd = gen_cli_golden.materialise(str(tmp_path))
    argv = [bam, fa, out, wd] + gold["flags"]
    cli.main_ctrl(cli.build_parser().parse_args(argv), argv, engine=EmulEngine())
    assert [l for l in open(out) if not l.startswith("##")] == gold["lines"]
    sys.modules.pop("pysam", None)


def _run_ties(engine, tmp_path, disable_tie_order=False, monkeypatch=None):
    bamio.build()
    gold = json.load(open(os.path.join(golden_util.GOLDEN, "cli_dataset2_ins_ties.json")))
    pk, fa, out, wd = gen_cli_golden.materialise(str(tmp_path), gold["seed"], gold["double_ins"])
    bam = _to_real_bam(pk, str(tmp_path / "real.bam"))
    if disable_tie_order:
        monkeypatch.setattr(cli, "ins_tie_swaps", lambda *a: [])
    argv = [bam, fa, out, wd] + gold["flags"]
    cli.main_ctrl(cli.build_parser().parse_args(argv), argv, engine=engine)
    import vcf_util
    return vcf_util.normalise_rnames([l for l in open(out) if not l.startswith("##")]), vcf_util.normalise_rnames(gold["lines"])


def test_cli_ins_ties_ordered_by_sequence_cpu(tmp_path, monkeypatch):
    """INS signatures that tie on (chr, int(pos), len, read) are ordered by their sequence strings (cuteSV:774): half of the
    INS-carrying reads report the insertion as two equal-length I ops at one position, -mi -1 keeps them apart.  The VCF body
    equals the REAL reference's; without the host-side tie ordering it does not (the golden pins which way ties fall)."""
    from emul_engine import EmulEngine
    lines, gold = _run_ties(EmulEngine(), tmp_path)
    assert len(lines) > 20 and lines == gold
    t2 = tmp_path / "off"
    t2.mkdir()
    lines_off, _ = _run_ties(EmulEngine(), t2, True, monkeypatch)
    assert lines_off != gold


@pytest.mark.gpu
def test_cli_ins_ties_ordered_by_sequence_gpu(engine, tmp_path):
    lines, gold = _run_ties(engine, tmp_path)
    assert lines == gold


def test_write_old_sigs_content(tmp_path):
    """--retain_work_dir --write_old_sigs: the legacy text dumps (cuteSV:766-816) hold exactly the de-duplicated signatures of
    the pickles, sorted by the reference's keys, one tab-separated line each."""
    import pickle as pk
    from emul_engine import EmulEngine
    lines, gold, wd = _run(EmulEngine(), tmp_path, 1, ["--retain_work_dir", "--write_old_sigs"])
    assert lines == gold
    idx = pk.load(open(os.path.join(wd, "sigindex.pickle"), "rb"))
    n_checked = 0
    for t, nf in (("DEL", 5), ("INS", 6), ("DUP", 5), ("INV", 6), ("TRA", 7)):
        rows = []
        with open(os.path.join(wd, t + ".pickle"), "rb") as f:
            for chrom in sorted(idx[t], key=lambda c: idx[t][c]):
                f.seek(idx[t][chrom])
                rows.extend(pk.load(f))
        text = [l.rstrip("\n").split("\t") for l in open(os.path.join(wd, t + ".sigs"))]
        assert len(text) == len(rows), t
        for e, l in zip(rows, text):
            assert len(l) == nf and l[0] == e[-2] and l[1] == e[-1], (t, l)
            if t in ("DEL", "DUP"):
                assert (int(l[2]), int(l[3]), l[4]) == (int(e[0]), int(e[1]), e[2])
            elif t == "INS":
                assert (int(l[2]), int(l[3]), l[4], l[5]) == (int(e[0]), int(e[1]), e[2], e[3])
            elif t == "INV":
                assert (l[2], int(l[3]), int(l[4]), l[5]) == (e[0], int(e[1]), int(e[2]), e[3])
            else:
                assert (l[2], int(l[3]), l[4], int(l[5]), l[6]) == (e[0], int(e[1]), e[2], int(e[3]), e[4])
            n_checked += 1
    assert n_checked > 200
