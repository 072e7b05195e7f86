"""CPU: the occupied-cell box used by the ray culling of the skipping mode is conservative with respect to the
reference's occupancy lookup (volume_renderer.py:992-1007, restated in oracle.nerf_oracle.is_empty_space), including
its clamping of points outside the [-2, 2]^3 volume."""
import torch

from oracle import nerf_oracle as O


def _box_contains(lo, hi, pts):
    inside = torch.ones(pts.shape[0], dtype=torch.bool)
    for c in range(3):
        inside &= (pts[:, c] >= lo[c]) & (pts[:, c] <= hi[c])
    return inside


def test_occupied_box_contains_every_point_with_an_occupied_lookup():
    from nerf_rep_for_test_b200.renderer import occupied_box
    g = torch.Generator().manual_seed(0)
    R = 32
    idx = torch.stack(torch.meshgrid([torch.arange(R)] * 3, indexing="ij"), -1)
    grids = []
    blob = torch.norm(idx.float() - torch.tensor([12.0, 20.0, 9.0]), dim=-1) <= 5.0       # interior blob
    grids.append(blob)
    grids.append(idx[..., 0] >= R - 2)                                                     # touches the +x boundary
    grids.append((idx[..., 1] == 0) & (idx[..., 2] > 10))                                  # touches the -y boundary
    grids.append(torch.rand(R, R, R, generator=g) < 0.001)                                 # sparse random cells
    one = torch.zeros(R, R, R, dtype=torch.bool); one[R - 1, 0, 17] = True                 # a single corner-ish cell
    grids.append(one)
    # points inside AND well outside the volume (lookups clamp)
    pts = (torch.rand(400000, 3, generator=g) - 0.5) * 6.0
    edge = torch.tensor([-2.0, 2.0, -2.0 + 4.0 / (R - 1), 2.0 - 4.0 / (R - 1), 0.0])
    pts = torch.cat([pts, torch.cartesian_prod(edge, edge, edge)])
    for grid in grids:
        lo, hi = occupied_box(grid)
        occupied = ~O.is_empty_space(grid, pts)
        assert int(occupied.sum()) > 0
        assert bool(_box_contains(lo, hi, pts[occupied]).all())
        # and the box is not trivially everything (for grids away from the boundary it is finite)
        if not bool(grid[0].any() | grid[-1].any() | grid[:, 0].any() | grid[:, -1].any() | grid[:, :, 0].any() | grid[:, :, -1].any()):
            assert all(abs(v) < 2.1 for v in lo + hi)
            assert float((~_box_contains(lo, hi, pts)).float().mean()) > 0.5


def test_occupied_box_of_an_empty_grid_culls_everything():
    from nerf_rep_for_test_b200.renderer import occupied_box
    lo, hi = occupied_box(torch.zeros(16, 16, 16, dtype=torch.bool))
    assert all(l > h for l, h in zip(lo, hi))
