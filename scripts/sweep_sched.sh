#!/bin/bash
# knob sweep of the scheduler on one GPU (value leg only); results -> gpurun_out/sweep_sched.jsonl
out=gpurun_out/sweep_sched.jsonl
mkdir -p gpurun_out
run() {  # env assignments..., then bench args after --
  local envs=() ; while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "# ${envs[*]} $*" >> $out
  env "${envs[@]}" timeout 600 python bench.py --value-only --steps 2 --warmup 1 "$@" >> $out 2>> gpurun_out/sweep_sched.err || echo "{\"failed\": \"${envs[*]} $*\"}" >> $out
}
run OH_X=0 --
run OH_ENC_PRIO=lo --
run OH_ENC_SMS=32 --
run OH_ENC_SMS=48 --
run OH_ENC_SMS=64 --
run OH_ENC_SMS=48 OH_CHECK=4 --
run OH_ENC_SMS=48 -- --slots 128
run OH_ENC_SMS=48 -- --slots 384
run OH_ASYNC_ENC=0 -- --blocking
cat $out
