import base64
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def load_golden():
    with open(os.path.join(HERE, "golden", "ref_vectors.json")) as f:
        doc = json.load(f)
    out = []
    for v in doc["vectors"]:
        out.append(dict(v, expr=base64.b64decode(v["expr"]), data=base64.b64decode(v["in"]),
                        out=base64.b64decode(v["out"])))
    return out
