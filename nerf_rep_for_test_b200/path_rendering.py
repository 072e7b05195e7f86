"""Novel-view path rendering around Renderer.render: the callers the reference keeps on its Renderer class
(src/models/nerf/renderer/volume_renderer.py:359-828) and that the evaluator invokes
(src/evaluators/nerf.py:605-640).  Same method names, arguments and return values, so the drop-in also serves those
call sites.  Not part of the hot path: camera-path geometry in numpy, one Renderer.render per pose, PNG / video
writing through OpenCV (imported lazily; the reference uses imageio for the video container, which this image
lacks).  Unlike the reference, a frame that fails to render raises instead of being replaced by a black frame
(volume_renderer.py:501-507)."""
import glob
import os

import numpy as np
import torch


def _cv2():
    try:
        import cv2
        return cv2
    except ImportError as e:          # pragma: no cover
        raise RuntimeError("OpenCV (cv2) is needed for PNG / video output of the novel-view helpers") from e


def _view_index(path):
    """evaluator / renderer file names are view0007_rgb.png, view0007_pred.png, ... (volume_renderer.py:611)."""
    return int(os.path.basename(path).split("_")[0].replace("view", ""))


class PathRenderingMixin:
    render_num = 120        # cfg.render_num
    fps = 30                # cfg.fps

    # ------------------------------------------------------------------ volume_renderer.py:359-428
    def generate_spiral_poses(self, poses, n_frames=None, n_rots=2, zrate=0.5):
        """[N,4,4] dataset poses -> [n_frames,4,4] spiral around their centroid (float64, like the reference):
        radius = mean distance to the centroid, n_rots turns, +-zrate of vertical swing, cameras look at the centre."""
        n_frames = int(n_frames or self.render_num)
        poses = poses.detach().cpu().numpy() if torch.is_tensor(poses) else np.asarray(poses)
        positions = poses[:, :3, 3]
        center = positions.mean(axis=0)
        unit = lambda v: v / np.linalg.norm(v)
        forward = unit(poses[:, :3, 2].mean(axis=0))
        up = unit(poses[:, :3, 1].mean(axis=0))
        right = unit(np.cross(forward, up))
        up = np.cross(right, forward)
        radius = np.linalg.norm(positions - center, axis=1).mean()
        i = np.arange(n_frames)
        theta = 2 * np.pi * n_rots * i / n_frames
        phi = zrate * np.sin(2 * np.pi * i / n_frames)
        cam_pos = center + radius * (np.cos(theta)[:, None] * right + np.sin(theta)[:, None] * forward) + phi[:, None] * up
        out = np.tile(np.eye(4), (n_frames, 1, 1))
        for k in range(n_frames):
            f = unit(center - cam_pos[k])
            r = unit(np.cross(f, up))
            out[k, :3, 0], out[k, :3, 1], out[k, :3, 2], out[k, :3, 3] = r, np.cross(r, f), f, cam_pos[k]
        return out

    def _default_intrinsics(self, hwf):
        H, W, focal = hwf
        return torch.tensor([[focal, 0, W / 2], [0, focal, H / 2], [0, 0, 1]], dtype=torch.float32, device=self.device)

    def _render_pose(self, pose, hwf, intrinsics):
        """one pose -> (rgb [H,W,3] in [0,1], disp [H,W] >= 0) on the device, clipped as :487-490 does."""
        H, W = int(hwf[0]), int(hwf[1])
        pose_t = torch.as_tensor(np.asarray(pose), dtype=torch.float32).to(self.device)
        ret = self.render({"pose": pose_t.unsqueeze(0), "intrinsics": intrinsics.to(self.device).unsqueeze(0), "H": H, "W": W})
        fine = "rgb_map" in ret
        rgb = ret["rgb_map" if fine else "rgb_map_0"].clamp(0, 1)
        disp = torch.nan_to_num(ret["disp_map" if fine else "disp_map_0"], nan=0.0, posinf=0.0).clamp_min(0)
        return rgb, disp

    # ------------------------------------------------------------------ volume_renderer.py:430-509
    @torch.no_grad()
    def render_path(self, render_poses, hwf, intrinsics=None, chunk_size=None):
        """-> (rgbs [N,H,W,3], disps [N,H,W]) float32 numpy.  `chunk_size` is accepted for signature parity and
        unused (the reference ignores it too)."""
        K = self._default_intrinsics(hwf) if intrinsics is None else intrinsics
        rgbs, disps = [], []
        for pose in render_poses:
            rgb, disp = self._render_pose(pose, hwf, K)
            rgbs.append(rgb)
            disps.append(disp)
        if not rgbs:
            return np.zeros((0, int(hwf[0]), int(hwf[1]), 3), np.float32), np.zeros((0, int(hwf[0]), int(hwf[1])), np.float32)
        return torch.stack(rgbs).cpu().numpy(), torch.stack(disps).cpu().numpy()     # one device->host copy each

    # ------------------------------------------------------------------ volume_renderer.py:511-616
    @torch.no_grad()
    def render_novel_view_sequence(self, poses, hwf, output_dir, exp_name, iteration=0, intrinsics=None, render_type="spiral"):
        """Renders the spiral (or the given poses), writes <output_dir>/novel_views/view%04d_{rgb,disp}.png and the
        video <output_dir>/<exp_name>_<render_type>_<iteration:06d>.mp4; returns (images_dir, video_path)."""
        cv2 = _cv2()
        if render_type == "spiral":
            render_poses = self.generate_spiral_poses(poses, n_frames=self.render_num)
        else:
            render_poses = poses.detach().cpu().numpy() if torch.is_tensor(poses) else np.asarray(poses)
        images_dir = os.path.join(output_dir, "novel_views")
        os.makedirs(images_dir, exist_ok=True)
        K = self._default_intrinsics(hwf) if intrinsics is None else intrinsics
        for i, pose in enumerate(render_poses):
            rgb, disp = self._render_pose(pose, hwf, K)
            rgb8 = (255 * rgb).to(torch.uint8).cpu().numpy()
            dmax = float(disp.max())
            disp8 = ((255 * disp / dmax) if dmax > 0 else disp).to(torch.uint8).cpu().numpy()
            cv2.imwrite(os.path.join(images_dir, "view%04d_rgb.png" % i), rgb8[..., ::-1])      # BGR for OpenCV
            cv2.imwrite(os.path.join(images_dir, "view%04d_disp.png" % i), disp8)
        video_path = os.path.join(output_dir, "%s_%s_%06d.mp4" % (exp_name, render_type, iteration))
        self.create_video_from_images(images_dir, video_path, pattern="*_rgb.png", sort_key=_view_index)
        return images_dir, video_path

    # ------------------------------------------------------------------ volume_renderer.py:618-707
    def _write_video(self, frames, output_video_path, fps):
        cv2 = _cv2()
        h, w = frames[0].shape[:2]
        os.makedirs(os.path.dirname(os.path.abspath(output_video_path)), exist_ok=True)
        vw = cv2.VideoWriter(output_video_path, cv2.VideoWriter_fourcc(*"mp4v"), float(fps), (w, h))
        if not vw.isOpened():
            raise RuntimeError("cannot open video writer for %s" % output_video_path)
        for f in frames:
            if f.shape[:2] != (h, w):
                f = cv2.resize(f, (w, h))
            vw.write(f)
        vw.release()

    def create_video_from_images(self, image_dir, output_video_path, fps=None, pattern="*.png", sort_key=None):
        """Frames = the images of `image_dir` matching `pattern`, ordered by `sort_key` (file name by default)."""
        cv2 = _cv2()
        files = sorted(glob.glob(os.path.join(image_dir, pattern)), key=sort_key)
        if not files:
            return None
        frames = [img for img in (cv2.imread(f) for f in files) if img is not None]
        if not frames:
            return None
        self._write_video(frames, output_video_path, fps or self.fps)
        return output_video_path

    # ------------------------------------------------------------------ volume_renderer.py:709-748
    def create_video_from_result_images(self, result_dir, output_video_path, image_type="pred", fps=None):
        images_dir = os.path.join(result_dir, "images")
        if not os.path.exists(images_dir):
            return None
        if image_type == "pred":
            return self.create_video_from_images(images_dir, output_video_path, fps=fps, pattern="*_pred.png", sort_key=_view_index)
        if image_type == "gt":
            return self.create_video_from_images(images_dir, output_video_path.replace(".mp4", "_gt.mp4"), fps=fps,
                                                 pattern="*_gt.png", sort_key=_view_index)
        if image_type == "both":
            return self.create_comparison_video(images_dir, output_video_path, fps=fps)
        raise ValueError("image_type must be 'pred', 'gt' or 'both'")

    # ------------------------------------------------------------------ volume_renderer.py:750-828
    def create_comparison_video(self, images_dir, output_video_path, fps=None):
        """prediction | ground truth side by side, for the views that have both files."""
        cv2 = _cv2()
        preds = {_view_index(f): f for f in glob.glob(os.path.join(images_dir, "*_pred.png"))}
        gts = {_view_index(f): f for f in glob.glob(os.path.join(images_dir, "*_gt.png"))}
        frames = []
        for k in sorted(set(preds) & set(gts)):
            p, g = cv2.imread(preds[k]), cv2.imread(gts[k])
            if p is None or g is None:
                continue
            if g.shape[:2] != p.shape[:2]:
                g = cv2.resize(g, (p.shape[1], p.shape[0]))
            frames.append(np.concatenate([p, g], axis=1))
        if not frames:
            return None
        self._write_video(frames, output_video_path, fps or self.fps)
        return output_video_path
