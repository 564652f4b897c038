#!/usr/bin/env python
"""tests/golden/orb_choice_golden.json: what the reference's own formatted_orb_choice
(bandupy/orbital_contributions.py:27-66) returns for the orbital selections `projected-unfold`
accepts.  Run in the build container (imports /root/reference with a stub for matplotlib, which
bandupy.constants touches at import time and which is not installed here)."""
import json
import sys
import types
from pathlib import Path

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/src/python_interface")
mpl = types.ModuleType("matplotlib")
mpl.get_backend = lambda: "agg"
mpl.use = lambda *a, **k: None
mpl.__version__ = "0.0"
sys.modules.setdefault("matplotlib", mpl)
from bandupy.orbital_contributions import formatted_orb_choice  # noqa: E402

spd = ['s', 'py', 'pz', 'px', 'dxy', 'dyz', 'dz2', 'dxz', 'dx2']
spdf = spd + ['fy(3x2-y2)', 'fxyz', 'fyz2', 'fz3', 'fxz2', 'fz(x2-y2)', 'fx(x2-3y2)']
choices = ['all', None, 's', 'p', 'd', 'f', 'sp2', 'spxy', 'spyx', 'spxz', 'spzx', 'spyz', 'spzy', 'sp3', 'px', ' Dz2 ',
           'P', ['s', 'd'], ['pz', 'px', 'py'], ['sp2', 'dxy'], ['d', 'p', 's'], ['f']]
cases = []
for supported in (spd, spdf):
    for ch in choices:
        cases.append(dict(choice=ch, supported=supported, picked=list(formatted_orb_choice(ch, list(supported)))))
Path(__file__).with_name("orb_choice_golden.json").write_text(json.dumps(cases, indent=0))
print(len(cases), "cases")
