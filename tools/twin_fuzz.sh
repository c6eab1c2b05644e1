#!/bin/bash
# Parity fuzz of the drop-in: the same C++ caller built against the reference's sources and against the drop-in
# (tests/cpp), 8 crowds x 5 level / crowd sizes, 200-400 frames each; every dump must be byte-identical.
# Run on the GPU box from the repo root:  bash tools/twin_fuzz.sh
cd tests/cpp/_bin
fail=0
for seed in 1 2 3 4 5 6 7 8; do
  for cfg in "8 6 300" "12 30 300" "16 3 400" "24 40 300" "10 100 200"; do
    ./driver_reference $cfg /tmp/r.bin $seed >/dev/null; ./driver_dropin $cfg /tmp/d.bin $seed >/dev/null
    if ! cmp -s /tmp/r.bin /tmp/d.bin; then echo "DIFF cfg=$cfg seed=$seed: $(cmp /tmp/r.bin /tmp/d.bin)"; fail=1; fi
  done
done
echo "fuzz done fail=$fail"
