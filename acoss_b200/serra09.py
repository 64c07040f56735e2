"""Drop-in ``Serra09`` plugin (reference: /root/reference/acoss/algorithms/rqa_serra09.py:18-83).

Same constructor arguments, attribute names and methods as the reference class.  What changes is
where the arithmetic runs: ``similarity(idxs)`` hands the whole index batch to the CUDA engine
(K1 OTI -> K2 CRP -> K3 Qmax) instead of calling essentia once per pair, and ``all_pairwise`` walks
the pair space in large tiles instead of one pair per call / joblib processes (a CUDA context must
not be forked under joblib; ``parallel`` / ``n_cores`` are accepted and ignored).
``normalize_by_length`` and ``getEvalStatistics`` keep the reference's results.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .algorithm_template import CoverAlgorithm
from .engine import Engine, pack_tracks

__all__ = ["Serra09"]


class Serra09(CoverAlgorithm):
    """
    Attributes (as in the reference): chroma_type, downsample_fac, all_feats, oti, kappa, tau, m.
    Extra keyword arguments: ``device`` (CUDA ordinal), ``features`` (in-memory feature dicts),
    ``gamma_o`` / ``gamma_e`` (essentia disOnset / disExtension, defaults 0.5), ``tile_pairs``.
    """

    def __init__(self, dataset_csv, datapath, chroma_type='hpcp', shortname='benchmark',
                 oti=True, kappa=0.095, tau=1, m=9, downsample_fac=40, device=0, features=None,
                 gamma_o=0.5, gamma_e=0.5, tile_pairs=1 << 16, cachedir="cache", engine=None,
                 _name="Serra09", _similarity_types=("main",)):
        self.oti = oti
        self.tau = tau
        self.m = m
        self.chroma_type = chroma_type
        self.kappa = kappa
        self.downsample_fac = downsample_fac
        self.gamma_o = gamma_o
        self.gamma_e = gamma_e
        self.tile_pairs = int(tile_pairs)
        self.device = device
        self.crp_path = _lib.CRP_AUTO              # ACOSS_CRP_AUTO (fast path) / ACOSS_CRP_EXACT
        self.all_feats = {}                      # cached (downsampled) chroma per song
        self._engine = engine
        self._resident = False
        CoverAlgorithm.__init__(self, dataset_csv=dataset_csv, name=_name, datapath=datapath,
                                shortname=shortname, cachedir=cachedir, features=features,
                                similarity_types=list(_similarity_types))

    # ------------------------------------------------------------------------------------------
    def load_features(self, i):
        """Downsampled chroma of song i, (n, 12) float32 (rqa_serra09.py:44-53).  With downsample_fac > 1 the
        per-bin block medians are computed on the GPU for the whole data set at once (acoss_set_tracks_raw) and
        mirrored into the host cache ``all_feats``; there is no host implementation of the aggregation."""
        if i not in self.all_feats:
            if int(self.downsample_fac) > 1:
                self.engine()                                    # fills all_feats for every song
            else:
                feats = CoverAlgorithm.load_features(self, i)
                self.all_feats[i] = np.array(feats[self.chroma_type])
        return self.all_feats[i]

    def params(self, **overrides):
        p = _lib.default_params(m=int(self.m), tau=int(self.tau), kappa=float(self.kappa), oti=int(bool(self.oti)),
                                gamma_o=float(self.gamma_o), gamma_e=float(self.gamma_e),
                                crp_path=int(self.crp_path))
        for k, v in overrides.items():
            setattr(p, k, v)
        return p

    def engine(self) -> Engine:
        """The CUDA engine with every song's frames resident in HBM (uploaded once)."""
        if self._engine is None:
            self._engine = Engine(self.device)
        if not self._resident:
            if int(self.downsample_fac) > 1:
                # the median aggregation of load_features runs on the GPU (k0_onramp.cu): raw frames go up once,
                # the downsampled tracks stay resident and are mirrored into the host cache all_feats
                if int(self.downsample_fac) > 128:
                    raise ValueError("downsample_fac > 128 is not supported by the GPU on-ramp (acoss_set_tracks_raw)")
                raws = []
                for i in range(self.N):
                    feats = CoverAlgorithm.load_features(self, i)         # also fills self.cliques
                    raws.append(np.asarray(feats[self.chroma_type], dtype=np.float32))
                frames, offsets = pack_tracks(raws)
                out_off = self._engine.set_tracks_raw(frames, offsets, int(self.downsample_fac))
                ds = self._engine.get_tracks()
                for i in range(self.N):
                    self.all_feats.setdefault(i, ds[out_off[i]:out_off[i + 1]])
            else:
                tracks = [np.asarray(self.load_features(i), dtype=np.float32) for i in range(self.N)]
                frames, offsets = pack_tracks(tracks)
                self._engine.set_tracks(frames, offsets)
            self._resident = True
        return self._engine

    def similarity(self, idxs):
        """Scores every (query i, reference j) row of idxs and stores it in Ds[key][i][j] for every
        key, like rqa_serra09.py:55-69 — in one batched GPU call."""
        idxs = np.asarray(idxs).reshape(-1, 2)
        if len(idxs) == 0:
            return
        scores = self.engine().score_pairs(idxs.astype(np.int32), self.params())
        for key in self.Ds.keys():
            self.Ds[key][idxs[:, 0], idxs[:, 1]] = scores

    def all_pairwise(self, parallel=0, n_cores=12, symmetric=False, precomputed=False):
        h5filename = "%s_Ds.h5" % self.get_cacheprefix()
        if precomputed:
            self.Ds = self._load_Ds(h5filename)
            self.get_all_clique_ids()
            return
        all_pairs = self._pair_array(symmetric)
        self.engine()                              # loads every song => cliques are complete
        for k0 in range(0, len(all_pairs), self.tile_pairs):
            self.similarity(all_pairs[k0:k0 + self.tile_pairs])
        if symmetric:
            for similarity_type in self.Ds:
                self.Ds[similarity_type] += self.Ds[similarity_type].T
        self._save_Ds(h5filename)

    def normalize_by_length(self):
        """Ds[i, j] /= sqrt(n_frames_j) (rqa_serra09.py:71-83): float64 divisor, float32 store."""
        fac = np.sqrt(np.array([self.load_features(j).shape[0] for j in range(self.N)], dtype=np.int64))
        for key in self.Ds.keys():
            D = np.asarray(self.Ds[key])
            self.Ds[key][:, :] = (D.astype(np.float64) / fac[None, :]).astype(np.float32)

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None
            self._resident = False
