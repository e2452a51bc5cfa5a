"""Console entry point mirroring the `-icl` flag of the reference's dispatcher
(wgdi/run.py:47 flag, :73-76 run_subprogram, :98 section name, :136/:150-152 `-icl ?`)."""
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
    parser = argparse.ArgumentParser(prog="wgdi_b200", description="B200-native `wgdi -icl` (improved collinearity)")
    parser.add_argument("-icl", dest="improvedcollinearity", metavar="CONF",
                        help="improved collinearity; `?` prints an example conf")
    args = parser.parse_args(argv)
    conf = args.improvedcollinearity
    if conf is None:
        parser.print_help()
        return 0
    if conf in ("?", "help", "example"):
        with open(os.path.join(HERE, "example", "collinearity.conf")) as f:
            print(f.read())
        return 0
    if not os.path.exists(conf):
        print(f"{conf} not exists")                       # wgdi/run.py:156-158
        return 0
    from .run_colliearity import mycollinearity
    run_subprogram(mycollinearity, conf, "collinearity")
    return 0


if __name__ == "__main__":
    sys.exit(main())
