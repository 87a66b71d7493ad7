cd $GRAFT_REPO_ROOT
for cfg in "1 24" "2 24" "2 12" "2 8" "2 48" "1 12"; do set -- $cfg
AVIF_B200_UP_STREAMS=$1 AVIF_B200_BAND_MB=$2 timeout 300 python bench.py --no-cpu-baseline --steps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('up_streams',$1,'band_mb',$2,'e2e',round(d['e2e']['value']))"
done
