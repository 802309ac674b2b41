"""Drop-in for the hot-path functions of the reference's cuteSV_genotype (cuteSV_genotype.py:10-173).

cal_GL is evaluated on the device (csv_cal_gl: special cases + rescale in the kernel, the libm part
from the host-built table).  The small host helpers used by the TRA genotyper are restated here."""
import numpy as np

from . import rows, runtime

Genotype = ["0/0", "0/1", "1/1"]


def cal_GL(c0, c1):
    """(GT, "PL0,PL1,PL2", GQ, QUAL) like cuteSV_genotype.py:33-56."""
    g = runtime.get_engine().cal_gl([int(c0)], [int(c1)])[0]
    qual = float(g["qual"])
    return Genotype[int(g["gt"])], "%d,%d,%d" % (int(g["pl"][0]), int(g["pl"][1]), int(g["pl"][2])), int(g["gq"]), (
        qual if (c0, c1) in ((3, 1), (6, 2)) else np.float64(qual))


def cal_GL_batch(c0, c1):
    return runtime.get_engine().cal_gl(c0, c1)


def cal_CIPOS(std, num):
    """cuteSV_genotype.py:58-60."""
    pos = int(1.96 * std / num ** 0.5)
    return "-%d,%d" % (pos, pos)


def threshold_ref_count(num):
    """cuteSV_genotype.py:62-70."""
    if num <= 2:
        return 20 * num
    if num <= 5:
        return 9 * num
    if num <= 15:
        return 7 * num
    return 5 * num


def count_coverage(chr, s, e, f, read_count, up_bound, itround):
    """cuteSV_genotype.py:72-93: BAM-order scan with early exit (used by the TRA genotyper only)."""
    status = 0
    iteration = 0
    primary_num = 0
    for i in f.fetch(chr, s, e):
        iteration += 1
        if i.flag not in (0, 16):
            continue
        primary_num += 1
        if i.reference_start < s and i.reference_end > e:
            read_count.add(i.query_name)
            if len(read_count) >= up_bound:
                status = 1
                break
        if iteration >= itround:
            status = 1 if float(primary_num / iteration) <= 0.2 else -1
            break
    return status


qual_str = rows.qual_str
