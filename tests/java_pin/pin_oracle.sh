#!/bin/bash
# tests/java_pin/pin_oracle.sh <checkout of rsutormin/KmerGutsJava> [workdir]
# One command that pins the repo's CPU oracle to the real Java on a box with a JDK (none exists in the build image):
# compiles the UNMODIFIED KmerGutsJava.java with tests/java_pin/GoldenDump.java, runs KmerGutsJava.main on the reference's own
# E. coli fixtures for the four flag sets (protein mode and 6-frame mode) and gatherHits on the 21 hand-traced FSM vectors,
# and compares everything with the oracle's output and with the SHA-256 values committed under tests/golden/.
set -euo pipefail
REF=${1:?usage: pin_oracle.sh <KmerGutsJava checkout> [workdir]}
W=${2:-/tmp/kg_pin}
HERE=$(cd "$(dirname "$0")/../.." && pwd)
cd "$HERE"
if ! command -v javac > /dev/null || ! command -v java > /dev/null; then
  # no JDK on this box: execute the same unmodified source through the mechanical transliterator instead (tests/java_pin/j2py.py)
  echo "no javac/java here: running KmerGutsJava.java through tests/java_pin/transliterated_pin.py"
  exec python tests/java_pin/transliterated_pin.py "$REF" "$W"
fi
python tests/java_pin/pin_oracle.py prepare "$W"
mkdir -p "$W/classes"
javac -nowarn -d "$W/classes" "$REF/lib/src/kmergutsjava/KmerGutsJava.java" tests/java_pin/GoldenDump.java
java -cp "$W/classes" kmergutsjava.GoldenDump kats "$W/kats.txt" "$W/java_kats.txt"
java -Xmx12g -cp "$W/classes" kmergutsjava.GoldenDump reports "$W/runs.txt"
python tests/java_pin/pin_oracle.py compare "$W"
# the Panama binding (JDK 22+): compile check only, no GPU needed
if javac --release 22 -d "$W/classes_ffm" kmergutsjava_b200/java/KmerGutsGpu.java 2> "$W/ffm.log"; then echo "PASS KmerGutsGpu.java compiles (JDK 22+ FFM)"; else echo "NOTE KmerGutsGpu.java did not compile here (needs JDK 22+): see $W/ffm.log"; fi
