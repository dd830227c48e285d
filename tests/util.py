import numpy as np


def maxrel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def l2rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def write_training_set(root, n=12, res=64, c_dim=5, depth=True, emb_dim=0, seed=0):
    """A training set in the layout the reference's dataset_tool.py writes (directory form): n RGB PNGs, dataset.json with integer labels and camera
    angles, optional 16-bit `<name>_depth.png` maps and an embeddings memmap + description (dataset.py:355-362).  Returns the dataset-config extras."""
    import json
    import os
    import PIL.Image
    rs = np.random.RandomState(seed)
    os.makedirs(root, exist_ok=True)
    labels, angles, rows = [], [], {}
    for i in range(n):
        name = f'{i % 2:05d}/img{i:08d}.png'
        os.makedirs(os.path.join(root, os.path.dirname(name)), exist_ok=True)
        PIL.Image.fromarray(rs.randint(0, 256, size=(res, res, 3)).astype(np.uint8), 'RGB').save(os.path.join(root, name))
        if depth:
            PIL.Image.fromarray(rs.randint(0, 65536, size=(res, res)).astype(np.uint16)).save(os.path.join(root, name[:-4] + '_depth.png'))
        labels.append([name, int(i % c_dim)])
        angles.append([name, [float(rs.uniform(-1.5, 1.5)), float(rs.uniform(0.8, 2.3)), 0.0]])
        rows[name] = n - 1 - i
    with open(os.path.join(root, 'dataset.json'), 'w') as f:
        json.dump(dict(labels=labels, camera_angles=angles), f)
    extra = dict(use_embeddings=False)
    if emb_dim:
        emb = rs.standard_normal((n, emb_dim)).astype(np.float32)
        epath, dpath = str(root) + '_emb.memmap', str(root) + '_emb.json'
        mm = np.memmap(epath, dtype='float32', mode='w+', shape=emb.shape); mm[:] = emb; mm.flush(); del mm
        with open(dpath, 'w') as f:
            json.dump(dict(shape=list(emb.shape), filepath_to_idx=rows), f)
        extra = dict(use_embeddings=True, embeddings_path=epath, embeddings_desc_path=dpath, _embeddings=emb, _rows=rows)
    return extra
