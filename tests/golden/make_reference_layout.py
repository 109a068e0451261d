"""Transcribes the layout of the reference's golden example output into a small fixture:
the banner and STATISTICS-table header lines of examples/refOutput/laplacian.txt (run in this
container, where /root/reference exists; the GPU box only reads the committed fixture)."""
import json
import os

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    lines = open(os.path.join(REF, "examples", "refOutput", "laplacian.txt")).read().splitlines()
    i = lines.index("STATISTICS SUMMARY:")
    header = lines[i:i + 6]                                   # title, blank, rule, two header rows, rule
    rows = [l for l in lines[i + 6:] if l.startswith("|")]
    setup = [l for l in lines if l.startswith(("Grid dimensions", "Processor topology", "Diffusion coeffs",
                                               "Discretization", "Number of solves"))]
    out = {"source": "examples/refOutput/laplacian.txt", "command": "laplacian -n 10 10 10 (verbosity 0x3, 5 solves)",
           "table_header": header, "table_rows": rows, "setup_lines": setup,
           "initial_res_norm": "1.00e+01", "cpu_default_iterations": 5}
    with open(os.path.join(HERE, "laplacian_layout.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote tests/golden/laplacian_layout.json")


if __name__ == "__main__":
    main()
