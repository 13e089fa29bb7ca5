import sys,json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "pcie", d["e2e"].get("pcie_h2d_gbs"), d["e2e"].get("pcie_d2h_gbs"), d["e2e"].get("pcie_bidir_gbs_each_way"), "cfg", d["config"]["group_pages"], d["config"]["lanes"])
    print("  iso", {k:v["us_per_page"] for k,v in d.get("stages_isolated",{}).items()})
