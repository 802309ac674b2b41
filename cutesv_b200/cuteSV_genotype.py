"""Drop-in for the hot-path entry point of the reference's cuteSV_genotype: cal_GL
(cuteSV_genotype.py:33-56), evaluated on the device (csv_cal_gl: special cases + rescale in the
kernel, the libm part from the host-built table).  overlap_cover / assign_gt / count_coverage have no
host counterpart here: they run inside csv_cluster."""
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


qual_str = rows.qual_str
