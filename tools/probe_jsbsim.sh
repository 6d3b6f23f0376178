#!/bin/bash
# Probe the GPU box for anything of the real JSBSim (VERDICT r1 "missing 1"): package, wheel, shared object,
# sources, a reference install, or an index to download it from. Output is committed under profiles/.
out=${1:-gpurun_out/r2_jsbsim_probe.txt}
{
echo "== date: $(date -u +%FT%TZ)  host: $(hostname)  nproc: $(nproc)"
echo "== python -c 'import jsbsim'"; python -c 'import jsbsim; print(jsbsim.__version__, jsbsim.__file__)' 2>&1 | tail -1
echo "== python -c 'import gymnasium'"; python -c 'import gymnasium; print(gymnasium.__version__)' 2>&1 | tail -1
echo "== python -c 'import stable_baselines3'"; python -c 'import stable_baselines3; print(stable_baselines3.__version__)' 2>&1 | tail -1
echo "== pip download jsbsim (no network expected)"; timeout 40 python -m pip download --no-deps -d /tmp/jsb_dl jsbsim 2>&1 | tail -3
echo "== pip install --no-index --find-links /opt/wheelhouse jsbsim"; timeout 40 python -m pip install --no-index --find-links /opt/wheelhouse --target /tmp/jsb_t jsbsim 2>&1 | tail -2
echo "== ls /opt/wheelhouse | grep -i -E 'jsb|gymnas|stable'"; ls /opt/wheelhouse 2>/dev/null | grep -i -E 'jsb|gymnas|stable' || echo "(none)"
echo "== find / -iname '*jsbsim*' (excluding this repo and /proc)"; find / -xdev \( -path /proc -o -path /sys -o -path "$PWD" -o -path /root/repo \) -prune -o -iname '*jsbsim*' -print 2>/dev/null | head -20; echo "(end of find)"
echo "== find / -iname 'FGFDMExec*' -o -iname 'libJSBSim*'"; find / -xdev \( -path /proc -o -path /sys \) -prune -o \( -iname 'FGFDMExec*' -o -iname 'libJSBSim*' \) -print 2>/dev/null | head; echo "(end of find)"
echo "== ls baseline/_ref"; ls -la baseline/_ref 2>&1 | head
echo "== ls /root/reference"; ls /root/reference 2>&1 | head -3
echo "== conda / apt caches"; ls /var/cache/apt/archives 2>/dev/null | grep -i jsb || echo "(no apt archive)"; which conda 2>&1 | tail -1
} > "$out" 2>&1
cat "$out"
