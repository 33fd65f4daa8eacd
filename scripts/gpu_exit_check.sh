# exit status of the GPU test process (the driver looks at it)
python -m pytest tests/test_host_cpp.py -m gpu -q -x -k "concurrent" > /tmp/p.log 2>&1; echo "concurrent rc=$?"; tail -2 /tmp/p.log
python -m pytest tests -x -q -m gpu > /tmp/p.log 2>&1; echo "pytest tests -x -q -m gpu rc=$?"; tail -3 /tmp/p.log
python -c "import __graft_entry__ as g; g.smoke()" > /tmp/s.log 2>&1; echo "smoke rc=$?"; tail -1 /tmp/s.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > /tmp/b.json 2>/tmp/b.err; echo "bench rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > /tmp/r.json 2>/tmp/r.err; echo "bench reference rc=$?"
