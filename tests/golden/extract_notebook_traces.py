"""Extract the verbose iteration tables embedded in the reference's notebooks into JSON fixtures.

Run in the build container (needs /root/reference):  python tests/golden/extract_notebook_traces.py
Sources (outputs of the unmodified reference package, produced by its authors):
  examples/acrobot/Acrobot.ipynb  cell 19   -> acrobot_notebook_trace.json
  examples/car/Car Escape.ipynb   cell 23   -> car_escape_notebook_trace.json
Each fixture keeps the printed numbers verbatim (as strings, so the printed precision is known).
"""
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ANSI = re.compile(r"\x1b\[[0-9;]*m")
NUM = r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|Inf|NaN)"


def cell_text(nb_path, cell):
    nb = json.load(open(nb_path))
    out = []
    for o in nb["cells"][cell].get("outputs", []):
        if "text" in o:
            out.append("".join(o["text"]))
    return ANSI.sub("", "".join(out))


def parse(text):
    """Rows are whitespace-separated numeric tables under a header line naming the columns."""
    tables, header = [], None
    for line in text.splitlines():
        s = line.strip()
        if not s or set(s) <= set("-_ "):
            continue
        toks = s.split()
        if all(re.fullmatch(NUM, t) for t in toks) and header is not None and len(toks) >= 3:
            tables[-1]["rows"].append(toks)
        elif any(t in ("iter", "cost", "c_max", "expected") for t in toks):
            header = toks
            if not tables or tables[-1]["columns"] != header or True:
                tables.append({"columns": header, "rows": []})
    return [t for t in tables if t["rows"]]


def main():
    for name, path, cell in (("acrobot_notebook_trace", "examples/acrobot/Acrobot.ipynb", 19),
                             ("car_escape_notebook_trace", "examples/car/Car Escape.ipynb", 23)):
        text = cell_text(os.path.join(REF, path), cell)
        tables = parse(text)
        json.dump({"source": "%s cell %d" % (path, cell), "tables": tables},
                  open(os.path.join(HERE, name + ".json"), "w"), indent=1)
        print(name, len(tables), "tables", sum(len(t["rows"]) for t in tables), "rows")


if __name__ == "__main__":
    main()
