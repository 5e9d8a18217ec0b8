#!/bin/bash
# TEST-ONLY: build the kernel sources for the host emulator (see cuda_emul.h).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="$HERE/../../pytorch-distributed-reinforcement-learning_b200/csrc"
OUT="$HERE/_build"
mkdir -p "$OUT"
OBJS=""
PIDS=""
for f in "$SRC"/*.cu; do
  b=$(basename "$f" .cu)
  if [ ! -f "$OUT/$b.o" ] || [ "$f" -nt "$OUT/$b.o" ] || [ -n "$(find "$SRC" "$HERE/cuda_emul.h" -newer "$OUT/$b.o" \( -name '*.cuh' -o -name '*.h' \) | head -1)" ]; then
    g++ -std=c++20 -O1 -g -fPIC -DB200RL_EMUL $B200RL_EMUL_DEFS -x c++ -c "$f" -o "$OUT/$b.o" -I"$SRC" -I"$HERE/../../include" -Wall -Wno-unknown-pragmas -Wno-unused-variable -Wno-unused-function -Wno-sign-compare &
    PIDS="$PIDS $!"
  fi
  OBJS="$OBJS $OUT/$b.o"
done
for p in $PIDS; do wait $p || { echo "emulator compile failed" >&2; exit 1; }; done
g++ -shared -o "$OUT/libb200rl_emul.so" $OBJS -lpthread
echo "built $OUT/libb200rl_emul.so"
