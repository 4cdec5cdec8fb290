#!/bin/bash
# The tutorial as shipped (Tutorial/Ginzburg-Landau, variables.yaml: k=10, q=1, Nw=42, w=0.05, TransientRemoval on)
# through the C host driver; prints the wall time.  Usage: tools/tutorial_run.sh <workdir>
set -e
W=${1:-/tmp/rsvdt_tutorial}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$W/results"
cp "$ROOT/tests/golden/A_GL" "$W/A_GL"
cat > "$W/variables.yaml" <<YAML
RootDir:            $W/
ResultsDir:         results/
OperatorDir:        A_GL
k:                  10
q:                  1
Nw:                 42
w:                  0.05
TwoPI:              false
TransientLength:    10
dt:                 0.003
TransientRemoval:   true
Display:            2
DiscFlg:            false
beta:               0.1
RandSeed:           14
SaveResultsOpt:     1
InputForcingFlg:    false
TransRun:           false
TransICDir:         none
YAML
cd "$W"
T0=$(date +%s%N)
"$ROOT/rsvd-delta-t_b200/RSVDt" -inputs variables.yaml | tail -25
T1=$(date +%s%N)
echo "tutorial full solve: $(( (T1 - T0) / 1000000 )) ms wall"
