"""`pb_render` with the reference's outputs (python/renderer.py:32-209), forward only (the backward runs inside
loss.total_loss / Engine.train_step)."""
from .engine import get_engine


def pb_render(camloc, raydir, color_gt, rnd, cos_anneal_ratio, conf):
    """Runs sample_points + pb_render and returns the reference's dictionary entries that exist as buffers:
    color_pixel (B,R,3), sdf_x_fg, grad_x_fg, alpha_fg, trans_fg, weights_fg, weights_bg, normal_pixel, ..."""
    eng = get_engine(conf)
    eng.train_step(camloc, raydir, color_gt, rnd, cos_anneal_ratio=cos_anneal_ratio, backward=False, keep=True)
    d = eng.debug
    B, R, N, Nb, M = d["dims"]
    NR, P, Df = B * R, B * R * N, eng.Df
    att = d["ATT"][:P]
    return dict(
        color_pixel=d["color"][:NR].view(B, R, 3), sdf_x_fg=d["sdf"][:P].view(B, R, N, 1),
        grad_x_fg=d["nrm"][:P].view(B, R, N, 3), alpha_fg=d["alpha_fg"][:P].view(B, R, N, 1),
        trans_fg=d["T"][:NR, :N].reshape(B, R, N, 1), weights_fg=d["w"][:NR, :N].reshape(B, R, N, 1),
        weights_bg=d["w"][:NR, N:].reshape(B, R, Nb, 1), normal_pixel=d["nhat"][:NR].view(B, R, 3),
        feature=d["O"][:P, :Df].reshape(B, R, N, Df), implicit=att[:, 0].reshape(B, R, N, 1),
        roughness=att[:, 1].reshape(B, R, N, 1), specular_reflectance=att[:, 2:5].reshape(B, R, N, 3),
        photogrammetric=att[:, 5].reshape(B, R, N, 1), x_fg=d["x_fg"].view(B, R, N, 3), mask=d["mask"])
