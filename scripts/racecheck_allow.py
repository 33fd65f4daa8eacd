"""Post-process a compute-sanitizer racecheck log: hazard reports grouped by the function(s) they name; anything
outside the allow-list fails (exit 1).  Allow-list = the shared-memory union-find of cc_tile_kernel (cc_find_s and
the compare-and-swap that links two roots): concurrent reads, path-halving writes and CAS on the parent array ARE
the lock-free algorithm -- whatever value a find reads is a valid ancestor.
    python scripts/racecheck_allow.py gpurun_out/sanitizer_racecheck.log"""
import collections, re, sys
ALLOW = ("cc_find_s", "cc_tile_kernel", "cc_union_s")
txt = open(sys.argv[1], errors="replace").read()
reports = re.findall(r"(?:Error|Warning): Race reported.*?(?=\n=========\s*\n|\Z)", txt, flags=re.S)
by = collections.Counter(); bad = collections.Counter()
for b in reports:
    names = sorted(set(re.findall(r"access at (?:[a-z ]+ )?(?:<unnamed>::)?([A-Za-z_0-9]+)[(<+]", b)))
    key = " / ".join(names) or "?"
    by[key] += 1
    if not names or not all(any(a in n for a in ALLOW) for n in names):
        bad[key] += 1
print("racecheck reports by function:", dict(by) or "none")
if bad:
    print("NOT on the allow-list:", dict(bad))
    sys.exit(1)
print("all reports on the allow-list" if by else "clean")
