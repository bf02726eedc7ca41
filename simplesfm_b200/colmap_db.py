"""COLMAP database writer for the matching path's outputs, without the `colmap` binary.

The reference's `ColmapBdManager` (`simple_sfm/colmap_utils/colmap_bd_utils.py:40-405`) obtains an empty
schema by running `colmap feature_extractor` (CPU SIFT over every image, `:249-273`; its own TODO at `:254` asks
for a binary-free path) and then rewrites three tables.  This module creates the COLMAP sqlite schema directly and
fills it with exactly the payloads the reference writes:

    images(image_id, name, camera_id = 1, prior_* = NULL)                      colmap_bd_utils.py:302-306
    keypoints(image_id, rows = K, cols = 4, data = float32 [x, y, 1, 0])        colmap_bd_utils.py:324-331
    matches(pair_id = 2147483647*min + max, rows = L, cols = 2, data = uint32)  colmap_bd_utils.py:86-91, 354-367
    match.txt  "name1 name2" per pair                                           colmap_bd_utils.py:371-372

Geometric verification (`colmap matches_importer`, `:374-381`) stays an external step: it is run only when a
`colmap` executable is on PATH.  Method names mirror the reference class so that the three database calls of
`generate_sparse_colmap` (`simple_sfm/dataset/simple_sfm_dataset_utils.py:129-139`) work unchanged; the COLMAP
subprocess steps the reference class also exposes (`mapper`, `triangulation`, `run_mapper`;
`colmap_bd_utils.py:94-147, 387-399`) are pass-throughs to the `colmap` executable and raise `FileNotFoundError`
with a clear message when it is not installed.

The shared camera (`--ImageReader.single_camera 1`, `:259`): when no size is given it is taken from the first
image of `images_folder_path`, like `colmap feature_extractor` does -- width, height and
f = 0.85 * max(w, h) (`default_focal_length_factor`, `:260`); with neither a size nor a readable image the
constructor raises instead of guessing.
"""
from __future__ import annotations

import logging
import os
import shutil
import sqlite3
import subprocess
import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

logger = logging.getLogger(__name__)

MAX_IMAGE_ID = 2147483647

_SCHEMA = (
    """CREATE TABLE IF NOT EXISTS cameras (
        camera_id INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, model INTEGER NOT NULL, width INTEGER NOT NULL,
        height INTEGER NOT NULL, params BLOB, prior_focal_length INTEGER NOT NULL)""",
    f"""CREATE TABLE IF NOT EXISTS images (
        image_id INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, name TEXT NOT NULL UNIQUE, camera_id INTEGER NOT NULL,
        prior_qw REAL, prior_qx REAL, prior_qy REAL, prior_qz REAL, prior_tx REAL, prior_ty REAL, prior_tz REAL,
        CONSTRAINT image_id_check CHECK(image_id >= 0 and image_id < {MAX_IMAGE_ID}),
        FOREIGN KEY(camera_id) REFERENCES cameras(camera_id))""",
    """CREATE TABLE IF NOT EXISTS keypoints (
        image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB,
        FOREIGN KEY(image_id) REFERENCES images(image_id) ON DELETE CASCADE)""",
    """CREATE TABLE IF NOT EXISTS descriptors (
        image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB,
        FOREIGN KEY(image_id) REFERENCES images(image_id) ON DELETE CASCADE)""",
    """CREATE TABLE IF NOT EXISTS matches (
        pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB)""",
    """CREATE TABLE IF NOT EXISTS two_view_geometries (
        pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB,
        config INTEGER NOT NULL, F BLOB, E BLOB, H BLOB, qvec BLOB, tvec BLOB)""",
    "CREATE UNIQUE INDEX IF NOT EXISTS index_name ON images(name)",
)

# COLMAP camera model ids (https://colmap.github.io/cameras.html)
CAMERA_MODELS = {"SIMPLE_PINHOLE": 0, "PINHOLE": 1, "SIMPLE_RADIAL": 2, "RADIAL": 3, "OPENCV": 4}


class ColmapDatabaseWriter(object):
    def __init__(self, db_dir: str, images_folder_path: Optional[str] = None, camera_type: Optional[str] = None,
                 camera_params: Optional[Sequence[float]] = None, camera_size: Optional[Tuple[int, int]] = None):
        self.db_dir = db_dir
        self.sparse_path = os.path.join(db_dir, 'sparse')
        self.dense_path = os.path.join(db_dir, 'dense')
        self.db_path = os.path.join(db_dir, 'database.db')
        self.images_folder_path = images_folder_path
        self.camera_type, self.camera_params = camera_type, camera_params
        os.makedirs(self.sparse_path, exist_ok=True)
        os.makedirs(self.dense_path, exist_ok=True)
        self.connection = sqlite3.connect(self.db_path)
        self.cursor = self.connection.cursor()
        for stmt in _SCHEMA:
            self.cursor.execute(stmt)
        self.connection.commit()
        if self.cursor.execute('SELECT COUNT(*) FROM cameras').fetchone()[0] == 0:
            # single shared camera, like `--ImageReader.single_camera 1` (colmap_bd_utils.py:259)
            w, h = camera_size if camera_size is not None else self._first_image_size(images_folder_path)
            if camera_type is not None and camera_params is not None:
                model, params, prior = CAMERA_MODELS[camera_type], np.asarray(camera_params, np.float64), 1
            else:   # default_focal_length_factor 0.85 (colmap_bd_utils.py:260), SIMPLE_RADIAL
                model, params, prior = 2, np.array([0.85 * max(w, h), w / 2.0, h / 2.0, 0.0], np.float64), 0
            self.cursor.execute('INSERT INTO cameras(camera_id, model, width, height, params, prior_focal_length) '
                                'VALUES(1, ?, ?, ?, ?, ?);', (model, int(w), int(h), params.tobytes(), prior))
            self.connection.commit()

    @staticmethod
    def _first_image_size(folder: Optional[str]) -> Tuple[int, int]:
        """(width, height) of the first image file in `folder` (name order), what `colmap feature_extractor`
        would record for the shared camera."""
        if folder is not None and os.path.isdir(folder):
            from PIL import Image
            for name in sorted(os.listdir(folder)):
                if name.lower().endswith(('.png', '.jpg', '.jpeg', '.bmp', '.tif', '.tiff', '.ppm', '.pgm')):
                    try:
                        with Image.open(os.path.join(folder, name)) as im:
                            return int(im.size[0]), int(im.size[1])
                    except OSError:
                        continue
        raise ValueError('ColmapDatabaseWriter: the camera size is unknown -- pass camera_size=(width, height) or an '
                         'images_folder_path that holds the images (the intrinsics written to the database decide '
                         'geometric verification and triangulation)')

    # ---- COLMAP subprocess steps of the reference class (colmap_bd_utils.py:94-147, 387-399): external binary
    @staticmethod
    def _colmap(args: List[str], check: bool = True):
        if shutil.which('colmap') is None:
            raise FileNotFoundError("the `colmap` executable is not on PATH: `colmap " + args[0] + "` is an external "
                                    "step of the reference pipeline and is not reimplemented here")
        subprocess.run(['colmap'] + args, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=check)

    @classmethod
    def mapper(cls, db_path, images_folder_path, sparse_path, camera_type=None, camera_params=None):
        args = ['mapper', '--Mapper.ba_refine_principal_point', '1', '--Mapper.filter_max_reproj_error', '2',
                '--Mapper.min_num_matches', '32', '--Mapper.extract_colors', '1', '--Mapper.max_num_models', '1',
                '--database_path', db_path, '--image_path', images_folder_path, '--output_path', sparse_path]
        if camera_type is not None and camera_params is not None:
            args += ['--Mapper.ba_refine_focal_length', '1', '--Mapper.ba_refine_extra_params', '1',
                     '--Mapper.ba_refine_principal_point', '1']
        cls._colmap(args)

    @classmethod
    def triangulation(cls, db_path, images_folder_path, sparse_path, output_path):
        cls._colmap(['point_triangulator', '--log_level', '1', '--Mapper.init_min_tri_angle', '4',
                     '--Mapper.init_min_num_inliers', '25', '--database_path', db_path,
                     '--image_path', images_folder_path, '--input_path', sparse_path, '--output_path', output_path])

    def run_mapper(self):
        self.mapper(self.db_path, self.images_folder_path, self.sparse_path, self.camera_type, self.camera_params)

    def close_bd(self):
        try:
            self.cursor.close()
            self.connection.close()
        except Exception:
            pass

    def __del__(self):
        self.close_bd()

    @staticmethod
    def image_ids_to_pair_id(image_id1, image_id2):
        if image_id1 > image_id2:
            return MAX_IMAGE_ID * image_id2 + image_id1
        return MAX_IMAGE_ID * image_id1 + image_id2

    def replace_images_data(self, images_names: Dict[int, str]):
        """Replaces current image data with new.  Indexes must start from 1."""
        self.cursor.execute('DELETE FROM images;')
        self.cursor.executemany(
            'INSERT INTO images(image_id, name, camera_id, prior_qw, prior_qx, prior_qy, prior_qz, '
            'prior_tx, prior_ty, prior_tz) VALUES(?, ?, 1, NULL, NULL, NULL, NULL, NULL, NULL, NULL);',
            [(int(i), str(n)) for i, n in images_names.items()])
        self.connection.commit()

    def replace_keypoints(self, images_names: Dict[int, str], frames: List):
        """Replaces current images keypoints with new ones (float32 [x, y, 1, 0] rows); clears descriptors."""
        self.cursor.execute('DELETE FROM keypoints;')
        self.cursor.execute('DELETE FROM descriptors;')
        rows = []
        for image_id_bd in images_names.keys():
            kp = frames[image_id_bd - 1].keypoints_poses.detach().cpu().numpy()
            data = np.concatenate([kp, np.ones((kp.shape[0], 1)), np.zeros((kp.shape[0], 1))], axis=1).astype(np.float32)
            rows.append((int(image_id_bd), data.shape[0], data.shape[1], data.tobytes()))
        self.cursor.executemany('INSERT INTO keypoints(image_id, rows, cols, data) VALUES(?, ?, ?, ?);', rows)
        self.connection.commit()

    def replace_and_verificate_matches(self, match_table: Dict[Tuple[int, int], np.ndarray],
                                       images_names: Dict[int, str], run_verification: Optional[bool] = None):
        logger.info('Starts matches importer.')
        start = time.time()
        self.cursor.execute('DELETE FROM matches;')
        txt, rows = [], []
        for (image_id1, image_id2), matches in match_table.items():
            txt.append((images_names[image_id1], images_names[image_id2]))
            if image_id1 > image_id2:
                matches = matches[:, [1, 0]]
            matches = np.ascontiguousarray(matches, dtype=np.uint32)
            rows.append((int(self.image_ids_to_pair_id(int(image_id1), int(image_id2))), matches.shape[0],
                         matches.shape[1], matches.tobytes()))
        self.cursor.executemany('INSERT INTO matches(pair_id, rows, cols, data) VALUES(?, ?, ?, ?);', rows)
        self.connection.commit()
        match_txt_path = os.path.join(self.db_dir, 'match.txt')
        with open(match_txt_path, 'w', encoding='utf-8') as f:
            for a, b in txt:
                f.write(f'{a} {b}\n')
        if run_verification is None:
            run_verification = shutil.which('colmap') is not None
        if run_verification:
            subprocess.run(['colmap', 'matches_importer', '--database_path', self.db_path, '--match_list_path',
                            match_txt_path, '--SiftMatching.use_gpu', '0', '--match_type', 'pairs'],
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        logger.info(f'Finished \nElapsed time: {float(time.time() - start)} \n')

    # ---- read-back helpers (tests, inspection)
    def read_keypoints(self, image_id: int) -> np.ndarray:
        r, c, blob = self.cursor.execute('SELECT rows, cols, data FROM keypoints WHERE image_id = ?;',
                                         (int(image_id),)).fetchone()
        return np.frombuffer(blob, np.float32).reshape(r, c)

    def read_matches(self, image_id1: int, image_id2: int) -> np.ndarray:
        pid = int(self.image_ids_to_pair_id(int(image_id1), int(image_id2)))
        r, c, blob = self.cursor.execute('SELECT rows, cols, data FROM matches WHERE pair_id = ?;', (pid,)).fetchone()
        return np.frombuffer(blob, np.uint32).reshape(r, c)
