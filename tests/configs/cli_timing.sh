#!/bin/bash
# tests/configs/cli_timing.sh -- where the wall-clock time of the command line goes on a FASTA of ~270k proteins (host parts vs GPU)
set -e
D=/tmp/kg_cli_timing; mkdir -p $D
python - <<PY
import sys, gzip
sys.path.insert(0, '.')
from tools import kg_synth as synth
synth.build_c0_fixture('tests/data/Ecoli_K12_W3110.faa.gz', '$D/KmerData')
ids, descr, seqs = synth.read_fasta_simple('tests/data/Ecoli_K12_W3110.faa.gz')
with open('$D/big.faa', 'wb') as f:
    for rep in range(20):
        for i, s in zip(ids, seqs):
            f.write(b'>%s_%d\n' % (i.encode(), rep))
            for a in range(0, len(s), 70):
                f.write(s[a:a+70] + b'\n')
PY
ls -la $D/big.faa
time kmergutsjava_b200/bin/kmer_guts_b200 -a -D $D/KmerData -q $D/big.faa -o $D/out.txt
wc -l $D/out.txt
time oracle/build/kmer_guts_oracle -a -D $D/KmerData -q $D/big.faa -o $D/out_oracle.txt -V direct
cmp $D/out.txt $D/out_oracle.txt && echo "reports identical"
