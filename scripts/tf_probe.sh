#!/bin/bash
# Probe: is the third-party layer that holds the hot path's arithmetic
# (ddsp==3.7.0 over TensorFlow, reference README.md:12-14) reachable on the GPU box?  Transcript ->
# gpurun_out/r02_tf_probe.txt (committed as profiles/r02_tf_probe.txt).
out=gpurun_out/r02_tf_probe.txt
{
  echo "== date: $(date -u +%FT%TZ)  host: $(hostname)"
  echo "== python: $(python --version 2>&1)"
  for m in tensorflow ddsp gin note_seq tensorflow_probability crepe; do
    echo "-- python -c 'import $m'"
    python -c "import $m; print('$m', getattr($m, '__version__', '?'))" 2>&1 | tail -1
  done
  echo "-- pip download ddsp==3.7.0 tensorflow-cpu (index)"
  timeout 60 python -m pip download --no-deps -d /tmp/tfprobe ddsp==3.7.0 tensorflow-cpu 2>&1 | tail -4
  echo "-- pip download from the offline wheelhouse"
  timeout 60 python -m pip download --no-index --find-links /opt/wheelhouse --no-deps -d /tmp/tfprobe ddsp tensorflow tensorflow-cpu 2>&1 | tail -3
  echo "-- ls /opt/wheelhouse | grep -i -e tensorflow -e ddsp -e gin"
  ls /opt/wheelhouse 2>/dev/null | grep -i -e tensorflow -e ddsp -e gin || echo "(none)"
  echo "-- ls baseline/_ref"
  ls baseline/_ref 2>&1 | head
  echo "-- find / -name 'tensorflow*' -maxdepth 6 (site-packages)"
  find / -maxdepth 6 \( -name 'tensorflow*' -o -name 'ddsp*' \) -not -path '/proc/*' -not -path '*/repo/*' -not -path '/tmp/*' 2>/dev/null | head
  echo "-- network: curl -sS -m 10 https://pypi.org/simple/ddsp/"
  timeout 15 curl -sS -m 10 -o /dev/null -w '%{http_code}\n' https://pypi.org/simple/ddsp/ 2>&1 | tail -1
  echo "-- /root/reference present on the box?"
  ls /root/reference 2>&1 | head -3
  echo "-- nproc / lscpu"
  nproc; lscpu | grep -e 'Model name' -e 'Socket' -e 'NUMA node(s)' -e '^CPU(s)'
  echo "-- nvidia-smi topo"
  nvidia-smi topo -m 2>&1 | head -20
} > $out 2>&1
cat $out
