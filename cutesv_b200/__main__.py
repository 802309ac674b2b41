"""`python -m cutesv_b200 <bam> <ref> <vcf> <work_dir> [flags]` -- same as `python -m cutesv_b200.cli`."""
import sys

from .cli import run

if __name__ == "__main__":
    sys.exit(run(sys.argv[1:]))
