"""Oracle: uniform t-sampling (a1) and query positions (a2).  Test infrastructure only."""
import torch as th


def get_intervals(t: th.Tensor, far: float):
    """reference barf/model_interpolation.py:114-132 (_get_intervals)."""
    t_end = th.zeros_like(t)
    t_end[:, :-1] = t[:, 1:].clone()
    t_end[:, -1] = far
    return t, t_end


def sample_uniform(near: float, far: float, batch: int, n_samples: int, jitter=None, offset_u=None,
                   offset_size: float = 0.0):
    """reference barf/model_interpolation.py:135-180 (_sample_t_stratified_uniform) with the
    random draws made explicit: jitter (B,S) ~ U[0,1) or None ("equidistant"), offset_u (B,1)."""
    interval_size = (far - near) / n_samples
    t = th.linspace(near, far - interval_size, n_samples).unsqueeze(0).repeat(batch, 1)
    if jitter is not None:
        t = t + jitter * interval_size
    if offset_size != 0 and offset_u is not None:
        t = t + offset_u.reshape(batch, 1) * interval_size * offset_size
    return get_intervals(t, far)


def t_query(t_start, t_end, strategy: str):
    """reference barf/model_interpolation.py:279-286."""
    if strategy == "left":
        return t_start
    if strategy == "middle":
        return (t_start + t_end) / 2
    raise ValueError(strategy)


def compute_positions(origins, directions, t_start, t_end, strategy: str):
    """reference barf/model_interpolation.py:288-312."""
    t = t_query(t_start, t_end, strategy)
    positions = origins.unsqueeze(1) + t.unsqueeze(2) * directions.unsqueeze(1)
    dirs = directions.unsqueeze(1).repeat(1, positions.shape[1], 1)
    return positions, dirs
