"""Seeded synthetic point clouds shaped like the reference's datasets.  No dataset is
available offline, so tests and bench.py use these.  Pure numpy, deterministic for a given seed."""
import numpy as np


def _surface_points(rng, n):
    """n points uniformly on the surface of a random union of 3-6 boxes / cylinders / spheres."""
    nprim = int(rng.integers(3, 7))
    kinds = rng.integers(0, 3, nprim)
    centres = rng.uniform(-1.0, 1.0, (nprim, 3))
    sizes = rng.uniform(0.2, 0.8, (nprim, 3))
    areas = np.empty(nprim)
    for i in range(nprim):
        a, b, c = sizes[i]
        if kinds[i] == 0:
            areas[i] = 2 * (a * b + b * c + a * c) * 4
        elif kinds[i] == 1:
            areas[i] = 2 * np.pi * a * (2 * c) + 2 * np.pi * a * a
        else:
            areas[i] = 4 * np.pi * a * a
    counts = rng.multinomial(n, areas / areas.sum())
    out = []
    for i in range(nprim):
        m = int(counts[i])
        if m == 0:
            continue
        a, b, c = sizes[i]
        if kinds[i] == 0:  # box: pick a face by area, then uniform on it
            fa = np.array([b * c, b * c, a * c, a * c, a * b, a * b])
            face = rng.choice(6, m, p=fa / fa.sum())
            u = rng.uniform(-1, 1, (m, 3)) * sizes[i]
            axis = face // 2
            sign = (face % 2) * 2.0 - 1.0
            u[np.arange(m), axis] = sign * sizes[i][axis]
            pts = u
        elif kinds[i] == 1:  # cylinder radius a, half height c
            side = 2 * np.pi * a * 2 * c
            cap = np.pi * a * a
            which = rng.choice(3, m, p=np.array([side, cap, cap]) / (side + 2 * cap))
            th = rng.uniform(0, 2 * np.pi, m)
            r = np.where(which == 0, a, a * np.sqrt(rng.uniform(0, 1, m)))
            z = np.where(which == 0, rng.uniform(-c, c, m), np.where(which == 1, c, -c))
            pts = np.stack([r * np.cos(th), r * np.sin(th), z], 1)
        else:  # sphere radius a
            v = rng.normal(size=(m, 3))
            pts = a * v / np.linalg.norm(v, axis=1, keepdims=True)
        out.append(pts + centres[i])
    pts = np.concatenate(out, 0)
    return pts[rng.permutation(pts.shape[0])]


def modelnet_like(n_clouds, n_points=8192, seed=11):
    """[n_clouds, n_points, 3] float32 in [0,1], scaled exactly like sample_modelnet.py:47-48."""
    rng = np.random.default_rng(seed)
    out = np.empty((n_clouds, n_points, 3), dtype=np.float32)
    for i in range(n_clouds):
        p = _surface_points(rng, n_points).astype(np.float32)
        p -= p.min()
        p /= p.max()
        out[i] = p
    return out


def uniform_cube(n_clouds, n_points, seed=11):
    return np.random.default_rng(seed).random((n_clouds, n_points, 3), dtype=np.float32)


def grid_quantised(n_clouds, n_points, depth=4, seed=11):
    """Points on the octree grid (i+1/2)/2^depth: exact distance ties are everywhere."""
    rng = np.random.default_rng(seed)
    cells = rng.integers(0, 2 ** depth, (n_clouds, n_points, 3))
    return ((cells + 0.5) / 2 ** depth).astype(np.float32)


def shapenet_like(n_clouds, n_points=2048, seed=11):
    """2048-point clouds shifted by +0.5 as sample_shapenet.py:161 does."""
    p = modelnet_like(n_clouds, n_points, seed) - 0.5
    return (p + 0.5).astype(np.float32)


def decompressed_like(clouds, sigma=2e-3, seed=12):
    """A 'decoded' cloud for eval: original + N(0, sigma) noise, points re-ordered."""
    rng = np.random.default_rng(seed)
    out = clouds + rng.normal(0, sigma, clouds.shape).astype(np.float32)
    for i in range(out.shape[0]):
        out[i] = out[i][rng.permutation(out.shape[1])]
    return out.astype(np.float32)


def scene_like(n_points=1_000_000, seed=11):
    """S3DIS-shaped room: 8 x 6 x 3 m, 6 planes + 20 boxes, surface-uniform + 1 cm noise, metres."""
    rng = np.random.default_rng(seed)
    room = np.array([8.0, 6.0, 3.0])
    n_planes = n_points // 2
    pts = []
    face = rng.integers(0, 6, n_planes)
    u = rng.uniform(0, 1, (n_planes, 3)) * room
    axis = face // 2
    u[np.arange(n_planes), axis] = (face % 2) * room[axis]
    pts.append(u)
    n_box = n_points - n_planes
    per = np.full(20, n_box // 20)
    per[: n_box - per.sum()] += 1
    for i in range(20):
        size = rng.uniform(0.2, 1.2, 3)
        c = rng.uniform(size, room - size)
        m = int(per[i])
        f = rng.integers(0, 6, m)
        v = rng.uniform(-1, 1, (m, 3)) * size
        ax = f // 2
        v[np.arange(m), ax] = ((f % 2) * 2.0 - 1.0) * size[ax]
        pts.append(v + c)
    p = np.concatenate(pts, 0) + rng.normal(0, 0.01, (n_points, 3))
    return p[rng.permutation(n_points)].astype(np.float32)[None]


def ae_shapes(k=128, d=16, L=7):
    """Parameter keys / shapes of the reference AE.AE(K, k, d, L) state_dict (/root/reference/AE.py:12-32)."""
    del L
    shapes = {}
    for i, (cin, cout) in enumerate([(3, 32), (32, 64), (64, 128)]):
        shapes[f"sa.conv{i}.weight"], shapes[f"sa.conv{i}.bias"] = (cout, cin, 1, 1), (cout,)
    for i, (cin, cout) in enumerate([(131, 128), (128, 256), (256, 512), (512, d)]):
        shapes[f"pn.mlp_Modules.{i}.0.weight"], shapes[f"pn.mlp_Modules.{i}.0.bias"] = (cout, cin, 1, 1), (cout,)
    for i, (cin, cout) in zip((0, 2, 4), [(d, 256), (256, 1024), (1024, k * 128)]):
        shapes[f"inv_pool.{i}.weight"], shapes[f"inv_pool.{i}.bias"] = (cout, cin), (cout,)
    for i, (cin, cout) in enumerate([(d + 128, 128), (128, 64), (64, 32), (32, 3)]):
        shapes[f"inv_mlp.mlp_Modules.{i}.0.weight"], shapes[f"inv_mlp.mlp_Modules.{i}.0.bias"] = (cout, cin, 1, 1), (cout,)
    return shapes


def seeded_state_dict(shapes, seed=11):
    """Deterministic PyTorch-default-like init (U(-1/sqrt(fan_in), 1/sqrt(fan_in))) for a dict of shapes, so tests,
    bench and the CPU baseline on any machine use identical weights without shipping a checkpoint."""
    import torch
    g = torch.Generator().manual_seed(seed)
    sd = {}
    fan_in = 1
    for key in sorted(shapes):
        shape = shapes[key]
        if key.endswith("weight"):
            fan_in = int(np.prod(shape[1:]))
        bound = 1.0 / np.sqrt(fan_in) if key.endswith(("weight", "bias")) else 1.0
        sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


def seeded_module_state(module, seed=11):
    """Deterministic parameters AND buffers for any nn.Module (BatchNorm running statistics included, so that eval-mode
    folding is exercised): same values on every machine, keyed only by the module's own state_dict shapes."""
    import torch
    g = torch.Generator().manual_seed(seed)
    sd = {}
    fan_in = 1
    for key, ref in module.state_dict().items():
        shape = tuple(ref.shape)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            sd[key] = torch.zeros(shape, dtype=ref.dtype)
        elif leaf == "running_var":
            sd[key] = torch.rand(shape, generator=g) + 0.5
        elif leaf == "running_mean":
            sd[key] = (torch.rand(shape, generator=g) - 0.5) * 0.2
        elif leaf == "weight" and len(shape) == 1:   # BatchNorm gamma
            sd[key] = torch.rand(shape, generator=g) + 0.5
        elif leaf == "weight":
            fan_in = int(np.prod(shape[1:]))
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) / np.sqrt(fan_in)
        else:                                        # biases / BatchNorm beta
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) / np.sqrt(max(fan_in, 1))
    return sd
