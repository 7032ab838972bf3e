"""`read_stone_info` -- mirror of utils/terrain_utils/terrain_utils.py:416-424."""
import numpy as np
import torch


def stone_info_from_array(arr, device='cuda:0'):
    """[S,6] (x,y,z,dx,dy,dz) -> f32 [S,7] on `device`; column 6 = max(dx,dy)/4 in the array's own dtype."""
    n = np.asarray(arr)
    radius = (np.maximum(n[:, 3], n[:, 4]) / 4).astype(np.float64).reshape(-1, 1)
    return torch.from_numpy(np.append(n, radius, axis=1)).to(device).float()


def read_stone_info(path, device='cuda:0'):
    return stone_info_from_array(np.load(path), device)
