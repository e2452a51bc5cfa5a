"""Console entry point mirroring the `-icl` and `-bi` flags of the reference's dispatcher
(wgdi/run.py:47 / :50 flags, :73-76 run_subprogram, :98 / :90 section names, :127 / :136 / :150-152 `-x ?`)."""
from __future__ import annotations

import argparse
import os
import sys

from . import base

HERE = os.path.dirname(os.path.abspath(__file__))


def run_subprogram(program, conf, name):
    options = base.load_conf(conf, name)
    r = program(options)
    r.run()


def main(argv=None):
    parser = argparse.ArgumentParser(prog="wgdi_b200", description="B200-native `wgdi -icl` (improved collinearity) and `-bi` (block information)")
    parser.add_argument("-icl", dest="improvedcollinearity", metavar="CONF",
                        help="improved collinearity; `?` prints an example conf")
    parser.add_argument("-bi", dest="blockinfo", metavar="CONF",
                        help="block information of a collinearity file; `?` prints an example conf")
    args = parser.parse_args(argv)
    tools = (("improvedcollinearity", "collinearity"), ("blockinfo", "blockinfo"))
    chosen = [(getattr(args, dest), section) for dest, section in tools if getattr(args, dest) is not None]
    if not chosen:
        parser.print_help()
        return 0
    for conf, section in chosen:
        if conf in ("?", "help", "example"):
            from .example import render
            print(render(section))
            continue
        if not os.path.exists(conf):
            print(f"{conf} not exists")                   # wgdi/run.py:156-158
            continue
        if section == "collinearity":
            from .run_colliearity import mycollinearity
            run_subprogram(mycollinearity, conf, section)
        else:
            from .block_info import block_info
            run_subprogram(block_info, conf, section)
    return 0


if __name__ == "__main__":
    sys.exit(main())
