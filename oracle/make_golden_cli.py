"""TEST INFRASTRUCTURE — extracts the command-line contract of the reference's entry point (flag names, defaults, types,
choices of every `parser.add_argument` in /root/reference/main_origin.py:80-139) into tests/golden/main_origin_flags.json,
so the drop-in `medvill_b200.main_origin` parser can be checked against it without the reference tree.
Run in the build container:   python oracle/make_golden_cli.py"""
import ast
import json
import os

REF = os.environ.get("MEDVILL_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "main_origin_flags.json")


def literal(node):
    try:
        return ast.literal_eval(node)
    except Exception:
        return ast.unparse(node)


def main():
    tree = ast.parse(open(os.path.join(REF, "main_origin.py")).read())
    flags = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "add_argument" and node.args:
            name = literal(node.args[0])
            kw = {k.arg: literal(k.value) for k in node.keywords if k.arg in ("default", "type", "choices", "nargs")}
            flags[name] = kw
    json.dump(flags, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote %d flags to %s" % (len(flags), OUT))


if __name__ == "__main__":
    main()
