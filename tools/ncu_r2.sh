# ncu --set full captures of the round-2 kernels (each after a plain run of the same command exited 0); keep the total under 64 MiB per call
set -x
prof() { # name skip cmd...
  name=$1; skip=$2; shift 2
  "$@" > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_ -s $skip -c 1 -o gpurun_out/r2_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"
}
for c in "$@"; do
  case $c in
    trail64) prof trail64 3 python tools/profile_case.py 2097152 64 bf16 none 4 trail --actions rng ;;
    trail64_lazy) prof trail64_lazy 3 python tools/profile_case.py 2097152 64 bf16 none 4 trail --actions rng --variant 4 ;;
    trail_obs64) prof trail_obs64 3 python tools/profile_case.py 131072 64 bf16 lut1 4 trail --actions rng ;;
    trail_obs64_popup3) prof trail_obs64_popup3 3 python tools/profile_case.py 65536 64 bf16 popup3 4 trail --actions rng ;;
    trail64_eps) prof trail64_eps 60 python tools/profile_case.py 2097152 64 bf16 none 4 trail --actions rng --policy free_eps --eps 0.1 --warmup 60 ;;
    bits_temper) prof bits_temper 4 python tools/profile_case.py 4194304 10 bf16 lut1 4 bits --slide temper --actions rng ;;
    bits10_lut1) prof bits10_lut1 4 python tools/profile_case.py 4194304 10 bf16 lut1 4 bits10 ;;
    bits10_popup3) prof bits10_popup3 4 python tools/profile_case.py 2097152 10 bf16 popup3 4 bits10 ;;
  esac
done
ls -la gpurun_out/*.ncu-rep
