"""Stand-in for the PyPI `plyfile` package, covering exactly what the reference uses: it is not
installed in this image and the reference neither vendors nor pins it.  Put this directory on PYTHONPATH only when
the real package is absent.

Call sites covered:  PlyData.read(path), PlyData([element]).write(path), PlyElement.describe(structured_array,
'vertex'), plydata['vertex'][name], plydata.elements[0][name], plydata.elements[0].properties[i].name
(LG/scene/dataset_readers.py:163-186, LG/scene/gaussian_model.py:239-314).  Format: PLY 1.0, binary_little_endian or
ascii, one or more elements of scalar properties (no list properties — the reference never writes any).
"""
import numpy as np

_PLY_TO_NP = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
              "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
              "double": "f8", "float64": "f8"}
_NP_TO_PLY = {"i1": "char", "u1": "uchar", "i2": "short", "u2": "ushort", "i4": "int", "u4": "uint", "f4": "float",
              "f8": "double"}


class PlyProperty:
    def __init__(self, name, dtype):
        self.name, self.dtype = name, np.dtype(dtype)

    def __repr__(self):
        return "PlyProperty(%r, %r)" % (self.name, self.dtype.str)


class PlyElement:
    def __init__(self, name, data):
        self.name, self.data = name, data
        self.properties = tuple(PlyProperty(n, data.dtype.fields[n][0]) for n in data.dtype.names)

    @staticmethod
    def describe(data, name):
        if data.dtype.names is None:
            raise ValueError("PlyElement.describe needs a structured array")
        return PlyElement(name, np.asarray(data))

    @property
    def count(self):
        return len(self.data)

    def __getitem__(self, key):
        return self.data[key]

    def __len__(self):
        return len(self.data)


class PlyData:
    def __init__(self, elements=(), text=False, byte_order="<"):
        self.elements = list(elements)
        self.text = text

    def __getitem__(self, name):
        for e in self.elements:
            if e.name == name:
                return e
        raise KeyError(name)

    def write(self, path):
        with open(path, "wb") as f:
            lines = ["ply", "format %s 1.0" % ("ascii" if self.text else "binary_little_endian")]
            for e in self.elements:
                lines.append("element %s %d" % (e.name, e.count))
                for p in e.properties:
                    lines.append("property %s %s" % (_NP_TO_PLY[p.dtype.str.lstrip("<>|=")], p.name))
            lines.append("end_header")
            f.write(("\n".join(lines) + "\n").encode("ascii"))
            for e in self.elements:
                if self.text:
                    for row in e.data:
                        f.write((" ".join(repr(v.item()) if hasattr(v, "item") else str(v) for v in row) + "\n").encode())
                else:
                    le = np.dtype([(n, e.data.dtype.fields[n][0].newbyteorder("<")) for n in e.data.dtype.names])
                    f.write(np.ascontiguousarray(e.data.astype(le)).tobytes())

    @staticmethod
    def read(path):
        with open(path, "rb") as f:
            if f.readline().strip() != b"ply":
                raise ValueError("%s: not a PLY file" % path)
            fmt, elements, cur = None, [], None
            while True:
                line = f.readline()
                if not line:
                    raise ValueError("%s: unterminated PLY header" % path)
                tok = line.decode("ascii").split()
                if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                    continue
                if tok[0] == "format":
                    fmt = tok[1]
                elif tok[0] == "element":
                    cur = [tok[1], int(tok[2]), []]
                    elements.append(cur)
                elif tok[0] == "property":
                    if tok[1] == "list":
                        raise NotImplementedError("list properties are not supported by this stand-in")
                    cur[2].append((tok[2], _PLY_TO_NP[tok[1]]))
                elif tok[0] == "end_header":
                    break
            out = []
            for name, count, props in elements:
                if fmt == "ascii":
                    rows = [f.readline().split() for _ in range(count)]
                    data = np.empty(count, dtype=[(n, t) for n, t in props])
                    for k, (n, t) in enumerate(props):
                        data[n] = np.array([r[k] for r in rows], dtype=np.float64).astype(t)
                else:
                    order = "<" if fmt == "binary_little_endian" else ">"
                    dt = np.dtype([(n, order + t if t not in ("i1", "u1") else t) for n, t in props])
                    data = np.frombuffer(f.read(dt.itemsize * count), dtype=dt, count=count)
                    data = data.astype(np.dtype([(n, t) for n, t in props]))
                out.append(PlyElement(name, data))
            return PlyData(out, text=(fmt == "ascii"))
