for cells in 12500 25000 50000; do for s in 0 1; do echo "== cells $cells solver $s"; CELLS=$cells SWNE_TC_SOLVER=$s python tools/prof_slots.py tc 2>&1 | grep "imi 50 profile False"; done; done
for s in 0 1; do for w in 0 1; do echo "== cfg1 solver $s wsolve_simt $w"; SWNE_TC_SOLVER=$s SWNE_TC_WSOLVE_SIMT=$w python tools/prof_cfg1.py 2>&1 | grep "^tc profile"; done; done
