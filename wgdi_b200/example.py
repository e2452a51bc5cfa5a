"""Example confs printed by `-icl ?` / `-bi ?` (the reference prints wgdi/example/collinearity.conf and
wgdi/example/blockinfo.conf, wgdi/run.py:127,136,150-152).  Kept as data: section name, then (key, value) with the
reference's spacing around '=' — including its misspelt `positon` key, which is what users copy."""

EXAMPLES = {
    "collinearity": [
        ('gff1 ', ' gff1 file'),
        ('gff2 ', ' gff2 file'),
        ('lens1 ', ' lens1 file'),
        ('lens2 ', ' lens2 file'),
        ('blast ', ' blast file'),
        ('blast_reverse ', ' false'),
        ('comparison ', ' genomes'),
        ('multiple  ', ' 1'),
        ('process ', ' 8'),
        ('evalue ', ' 1e-5'),
        ('score ', ' 100'),
        ('grading ', ' 50,30,25'),
        ('mg ', ' 25,25'),
        ('pvalue ', ' 1'),
        ('repeat_number ', ' 20'),
        ('positon ', ' order'),
        ('savefile ', ' collinearity file'),
    ],
    "blockinfo": [
        ('blast ', ' blast file'),
        ('gff1 ', '  gff1 file'),
        ('gff2 ', '  gff2 file'),
        ('lens1 ', ' lens1 file'),
        ('lens2 ', ' lens2 file'),
        ('collinearity ', ' collinearity file'),
        ('score ', ' 100'),
        ('evalue ', ' 1e-5'),
        ('repeat_number ', ' 20'),
        ('position ', ' order'),
        ('ks ', ' ks file'),
        ('ks_col ', ' ks_NG86'),
        ('savefile ', ' block information (*.csv)'),
    ],
}


def render(section):
    return "\n".join(["[%s]" % section] + [k + "=" + v for k, v in EXAMPLES[section]]) + "\n"
