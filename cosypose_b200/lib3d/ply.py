"""Minimal PLY reader for the object models of the path (BOP `models/obj_%06d.ply`, the meshes the
reference turns into URDFs for its renderer: cosypose/scripts/convert_models_to_urdf.py, and samples its
point clouds from: lib3d/rigid_mesh_database.py:84-91 via trimesh).  Handles what those files use:
ascii / binary_little_endian / binary_big_endian, a `vertex` element with scalar properties
(x y z, optional nx ny nz, red green blue [alpha], texture_u texture_v) and a `face` element whose
first list property holds the vertex indices; polygons are fan-triangulated.
"""
import numpy as np

_TYPES = {'char': 'i1', 'int8': 'i1', 'uchar': 'u1', 'uint8': 'u1', 'short': 'i2', 'int16': 'i2',
          'ushort': 'u2', 'uint16': 'u2', 'int': 'i4', 'int32': 'i4', 'uint': 'u4', 'uint32': 'u4',
          'float': 'f4', 'float32': 'f4', 'double': 'f8', 'float64': 'f8'}


def read_ply(path):
    """-> dict(vertices [Nv,3] float32, faces [Nf,3] int32, colors [Nv,3] float32 in [0,1] or None)."""
    with open(path, 'rb') as f:
        data = f.read()
    end = data.find(b'end_header')
    assert data[:3] == b'ply' and end > 0, f'{path}: not a PLY file'
    header = data[:end].decode('ascii', 'replace').splitlines()
    body = data[data.index(b'\n', end) + 1:]
    fmt, elements = None, []
    for line in header:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == 'format':
            fmt = tok[1]
        elif tok[0] == 'element':
            elements.append(dict(name=tok[1], count=int(tok[2]), props=[]))
        elif tok[0] == 'property':
            if tok[1] == 'list':
                elements[-1]['props'].append(dict(name=tok[4], list=(_TYPES[tok[2]], _TYPES[tok[3]])))
            else:
                elements[-1]['props'].append(dict(name=tok[2], dtype=_TYPES[tok[1]]))
    assert fmt in ('ascii', 'binary_little_endian', 'binary_big_endian'), f'{path}: format {fmt}'
    out = {}
    if fmt == 'ascii':
        tokens = body.split()
        pos = 0
        for el in elements:
            if all('dtype' in p for p in el['props']):
                n = el['count'] * len(el['props'])
                tab = np.asarray(tokens[pos:pos + n], dtype=np.float64).reshape(el['count'], len(el['props']))
                pos += n
                out[el['name']] = {p['name']: tab[:, i] for i, p in enumerate(el['props'])}
            else:
                rows = []
                for _ in range(el['count']):
                    row = {}
                    for p in el['props']:
                        if 'list' in p:
                            k = int(tokens[pos])
                            row[p['name']] = [int(float(t)) for t in tokens[pos + 1:pos + 1 + k]]
                            pos += 1 + k
                        else:
                            row[p['name']] = float(tokens[pos])
                            pos += 1
                    rows.append(row)
                out[el['name']] = rows
    else:
        e = '<' if fmt == 'binary_little_endian' else '>'
        pos = 0
        for el in elements:
            if all('dtype' in p for p in el['props']):
                dt = np.dtype([(p['name'], e + p['dtype']) for p in el['props']])
                tab = np.frombuffer(body, dtype=dt, count=el['count'], offset=pos)
                pos += dt.itemsize * el['count']
                out[el['name']] = {p['name']: tab[p['name']] for p in el['props']}
            else:
                rows = _read_list_element(body, pos, el, e)
                pos = rows.pop()
                out[el['name']] = rows
    v = out['vertex']
    vertices = np.stack([np.asarray(v[k], dtype=np.float32) for k in ('x', 'y', 'z')], axis=1)
    colors = None
    if all(k in v for k in ('red', 'green', 'blue')):
        colors = np.stack([np.asarray(v[k], dtype=np.float32) for k in ('red', 'green', 'blue')], axis=1)
        if colors.size and colors.max() > 1.0:
            colors = colors / np.float32(255)
    faces = []
    for row in out.get('face', []):
        ids = next(val for val in row.values() if isinstance(val, (list, np.ndarray)))
        for k in range(1, len(ids) - 1):
            faces.append((ids[0], ids[k], ids[k + 1]))
    faces = np.asarray(faces, dtype=np.int32).reshape(-1, 3)
    return dict(vertices=vertices, faces=faces, colors=colors)


def _read_list_element(body, pos, el, e):
    """Rows of an element with list properties; the fast path covers the usual `uchar 3 + 3 ints` triangle table.
    Returns the rows with the end offset appended."""
    props = el['props']
    if len(props) == 1 and 'list' in props[0] and el['count'] > 0:
        ct, it = props[0]['list']
        k = int(np.frombuffer(body, dtype=e + ct, count=1, offset=pos)[0])
        rec = np.dtype([('n', e + ct), ('ids', e + it, (k,))])
        if pos + rec.itemsize * el['count'] <= len(body):
            tab = np.frombuffer(body, dtype=rec, count=el['count'], offset=pos)
            if (tab['n'] == k).all():
                rows = [{props[0]['name']: ids} for ids in tab['ids'].astype(np.int64)]
                rows.append(pos + rec.itemsize * el['count'])
                return rows
    rows = []
    for _ in range(el['count']):
        row = {}
        for p in props:
            if 'list' in p:
                ct, it = p['list']
                k = int(np.frombuffer(body, dtype=e + ct, count=1, offset=pos)[0])
                pos += np.dtype(ct).itemsize
                row[p['name']] = np.frombuffer(body, dtype=e + it, count=k, offset=pos).astype(np.int64).tolist()
                pos += np.dtype(it).itemsize * k
            else:
                row[p['name']] = float(np.frombuffer(body, dtype=e + p['dtype'], count=1, offset=pos)[0])
                pos += np.dtype(p['dtype']).itemsize
        rows.append(row)
    rows.append(pos)
    return rows
