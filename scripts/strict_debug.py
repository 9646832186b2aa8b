import gzip, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from posidonius_b200.case import case_from_dict
from posidonius_b200.ensemble import Ensemble
G = os.path.join(ROOT, "tests", "golden")
man = json.load(open(os.path.join(G, "manifest.json")))
arith = int(sys.argv[1])
nsys = int(sys.argv[2])
names = sys.argv[3:] or sorted(man["fixtures"])
for name in names:
    if name.startswith("c"):
        d = json.load(gzip.open(os.path.join(G, "configs", name + ".json.gz"), "rt"))
    else:
        d = json.load(gzip.open(os.path.join(G, man["fixtures"][name]["case"]), "rt"))
    case, tables = case_from_dict(d)
    if arith and case.general_relativity_implementation in (1, 2) and case.consider_general_relativity:
        continue
    ens = Ensemble(case, tables, n_systems=nsys, arithmetic=arith)
    ens.initialize_physical_values()
    ens.iterate(5)
    print(name, "ok", ens.status()[0][:4])
    ens.close()
