for v in "0 0 ab" "1 0 ab" "1 0 ba"; do set -- $v; E2E_SPANS=1 E2E_ORDER=$3 SB_STREAM_CLASSIFY=$1 SB_STREAM_CTAS=$2 python scripts/e2e_quick.py 10 2>&1 | tail -22; done
