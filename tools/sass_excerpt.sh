#!/bin/bash
# profiles/r02_sass_k_env_step_obs_rt_and_staged.txt: the TMA / mbarrier / L2-policy / packed-fp32 lines of the two dominant kernels
out=${1:-profiles/r02_sass_k_env_step_obs_rt_and_staged.txt}
{
echo "# SASS evidence (cuobjdump -sass pm-rl_b200/libpmrl_b200.so, sm_100a cubins only) — TMA bulk copies (UBLKCP), mbarrier waits (SYNCS),"
echo "# L2 prefetch (CCTL.E.PF2), no-L1-allocate 16-byte loads (LDG.E.NA.128), packed fp32 (FADD2/FMUL2/FFMA2), 3-input min/max (FMNMX3)."
echo "# Regenerate: tools/sass_excerpt.sh"
for fn in '_ZN4pmrl17k_env_step_obs_rtILi4ELb0ELi4ELi50ELi1ELb0EEEvNS_10StepParamsE' '_ZN4pmrl17k_env_step_obs_rtILi4ELb0ELi4ELi50ELi2ELb0EEEvNS_10StepParamsE' '_ZN4pmrl17k_env_step_obs_rtILi2ELb0ELi2ELi50ELi1ELb0EEEvNS_10StepParamsE' '_ZN4pmrl17k_env_step_stagedILi16ELb1ELb1ELi8ELi2ELb1EEEvNS_10StepParamsE'; do
  echo; echo "## $(echo $fn | c++filt)"
  cuobjdump -sass -fun "$fn" pm-rl_b200/libpmrl_b200.so > /tmp/k.sass 2>/dev/null
  echo "arch: $(grep -m1 'arch =' /tmp/k.sass)"
  echo "instruction count: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' /tmp/k.sass)"
  echo "opcode counts: UBLKCP=$(grep -c UBLKCP /tmp/k.sass) SYNCS=$(grep -c 'SYNCS' /tmp/k.sass) FENCE.VIEW.ASYNC=$(grep -c 'FENCE.VIEW.ASYNC' /tmp/k.sass) UTMALDG=$(grep -c UTMALDG /tmp/k.sass) LDS.128=$(grep -c 'LDS.128' /tmp/k.sass) LDG.E.128=$(grep -c 'LDG.E.*128' /tmp/k.sass) STG.E.128=$(grep -c 'STG.E.128' /tmp/k.sass) FADD2=$(grep -c FADD2 /tmp/k.sass) FMUL2=$(grep -c FMUL2 /tmp/k.sass) FFMA2=$(grep -c FFMA2 /tmp/k.sass) FMNMX3=$(grep -c FMNMX3 /tmp/k.sass) SHFL=$(grep -c SHFL /tmp/k.sass) HMMA/UTCMMA=$(grep -cE 'HMMA|UTC.*MMA' /tmp/k.sass)"
  grep -E "UBLKCP|SYNCS|FENCE.VIEW.ASYNC|UTMALDG|LDG\.E.*128|LDS\.128|CCTL|ATOMG" /tmp/k.sass | sed 's/^\s*//' | cut -c1-150 | head -48
done; } > $out
