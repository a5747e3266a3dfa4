#!/bin/bash
# Which Blackwell-specific instructions the built library contains, per kernel: tcgen05 MMA / commit / TMEM loads, bulk
# asynchronous copies (cp.async.bulk) and their mbarrier transaction counts, programmatic dependent launch, vector
# reductions.  Runs without a GPU (cuobjdump on the in-tree .so).   usage: scripts/sass_mnemonics.sh > profiles/<file>
LIB=${1:-mvin_b200/lib/libmvin_b200.so}
echo "# cuobjdump -sass $LIB ($(stat -c %s $LIB) bytes), $(nvcc --version | tail -2 | head -1)"
cuobjdump -sass $LIB | awk '
  /Function :/ { fn = $3 }
  {
    if ($0 ~ /UTCHMMA/) c[fn, "UTCHMMA"]++;            # tcgen05.mma
    if ($0 ~ /UTCBAR/) c[fn, "UTCBAR"]++;              # tcgen05.commit -> mbarrier
    if ($0 ~ /LDTM/) c[fn, "LDTM"]++;                  # tcgen05.ld (TMEM -> registers)
    if ($0 ~ /UTCATOMSWS|UTCALLOC/) c[fn, "TMEM alloc"]++;
    if ($0 ~ /UBLKCP/) c[fn, "UBLKCP"]++;              # cp.async.bulk
    if ($0 ~ /UBLKPF/) c[fn, "UBLKPF"]++;              # cp.async.bulk.prefetch
    if ($0 ~ /SYNCS.*TRANS/) c[fn, "SYNCS.TRANS"]++;   # mbarrier expect_tx / complete_tx
    if ($0 ~ /ACQBULK/) c[fn, "ACQBULK"]++;            # griddepcontrol.wait
    if ($0 ~ /PREEXIT/) c[fn, "PREEXIT"]++;            # griddepcontrol.launch_dependents
    if ($0 ~ /REDG\.E\.ADD\.F32x4/) c[fn, "RED.v4"]++;              # red.global.add.v4.f32
    if ($0 ~ /HMMA\.1688\.F32\.TF32/) c[fn, "HMMA.TF32"]++;
    if ($0 ~ /FFMA2|FADD2/) c[fn, "FFMA2/FADD2"]++;
    seen[fn] = 1
  }
  END {
    n = split("UTCHMMA UTCBAR LDTM UBLKCP UBLKPF SYNCS.TRANS ACQBULK PREEXIT RED.v4 HMMA.TF32 FFMA2/FADD2", cols, " ");
    for (i = 1; i <= n; i++) tot[cols[i]] = 0;
    for (f in seen) for (i = 1; i <= n; i++) tot[cols[i]] += c[f, cols[i]];
    printf "## totals over %d kernels\n", length(seen);
    for (i = 1; i <= n; i++) printf "%-12s %d\n", cols[i], tot[cols[i]];
    printf "\n## kernels with tcgen05 or bulk-copy instructions (counts)\n";
    for (f in seen) if (c[f, "UTCHMMA"] + c[f, "UBLKCP"] + c[f, "UBLKPF"] > 0)
      printf "%s  UTCHMMA=%d UTCBAR=%d LDTM=%d UBLKCP=%d UBLKPF=%d SYNCS.TRANS=%d\n", f, c[f, "UTCHMMA"], c[f, "UTCBAR"], c[f, "LDTM"], c[f, "UBLKCP"], c[f, "UBLKPF"], c[f, "SYNCS.TRANS"];
  }' | (sed -u '/^## kernels/q'; sort) | while read -r line; do
    case "$line" in _Z*) echo "$(echo "${line%% *}" | c++filt | cut -c1-110)  ${line#* }";; *) echo "$line";; esac
  done
