"""Boundary data types of the path: named tensors sharing a leading dimension, optionally
paired with a pandas `infos` table (one row per entry).

Interface-compatible with the reference containers the predictors exchange
(reference: cosypose/utils/tensor_collection.py:7-19 `concatenate`, :22-102 `TensorCollection`,
:105-174 `PandasTensorCollection`): attribute access to tensors, integer / array indexing,
`.to/.cuda/.cpu/.float`, `clone`, `merge_df`, pickling, and `gather_distributed`.  The
distributed gather uses one `all_gather_object` instead of the reference's per-rank files on a
shared filesystem (tensor_collection.py:142-163).
"""
import pandas as pd
import torch


class TensorCollection:
    def __init__(self, **tensors):
        object.__setattr__(self, '_tensors', {})
        for name, value in tensors.items():
            self.register_tensor(name, value)

    # -- registry -----------------------------------------------------------------------------
    def register_tensor(self, name, tensor):
        self._tensors[name] = tensor

    def delete_tensor(self, name):
        del self._tensors[name]

    @property
    def tensors(self):
        return self._tensors

    @property
    def device(self):
        return next(iter(self._tensors.values())).device

    def __getattr__(self, name):
        # only reached when normal lookup fails
        tensors = self.__dict__.get('_tensors')
        if tensors is not None and name in tensors:
            return tensors[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if '_tensors' not in self.__dict__:
            raise ValueError('Please call __init__')
        if name in self._tensors:
            self._tensors[name] = value
        else:
            object.__setattr__(self, name, value)

    def _describe(self):
        return ''.join(f'    {k}: {t.shape} {t.dtype} {t.device},\n' for k, t in self._tensors.items())

    def __repr__(self):
        return f'{type(self).__name__}(\n{self._describe()})'

    # -- indexing / conversion ----------------------------------------------------------------
    def _index_tensors(self, ids):
        return {k: t[ids] for k, t in self._tensors.items()}

    def __getitem__(self, ids):
        return TensorCollection(**self._index_tensors(ids))

    def to(self, target):
        for k in list(self._tensors):
            self._tensors[k] = self._tensors[k].to(target)
        return self

    def cuda(self):
        return self.to('cuda')

    def cpu(self):
        return self.to('cpu')

    def float(self):
        return self.to(torch.float)

    def double(self):
        return self.to(torch.double)

    def half(self):
        return self.to(torch.half)

    def clone(self):
        return TensorCollection(**{k: t.clone() for k, t in self._tensors.items()})

    def __getstate__(self):
        return {'tensors': self._tensors}

    def __setstate__(self, state):
        TensorCollection.__init__(self, **state['tensors'])


class PandasTensorCollection(TensorCollection):
    def __init__(self, infos, **tensors):
        super().__init__(**tensors)
        self.infos = infos.reset_index(drop=True)
        self.meta = dict()

    def __len__(self):
        return len(self.infos)

    def __repr__(self):
        return (f'{type(self).__name__}(\n{self._describe()}{"-" * 40}\n'
                f'    infos:\n{self.infos!r}\n)')

    def __getitem__(self, ids):
        infos = self.infos.iloc[ids].reset_index(drop=True)
        return PandasTensorCollection(infos, **self._index_tensors(ids))

    def merge_df(self, df, *args, **kwargs):
        infos = self.infos.merge(df, how='left', *args, **kwargs)
        assert len(infos) == len(self.infos)
        assert (infos.index == self.infos.index).all()
        return PandasTensorCollection(infos=infos, **self.tensors)

    def clone(self):
        return PandasTensorCollection(self.infos.copy(), **{k: t.clone() for k, t in self._tensors.items()})

    def gather_distributed(self, tmp_dir=None):
        """Concatenation of every rank's collection, in rank order, returned on every rank
        (the reference returns it on rank 0 only).  `tmp_dir` is accepted and ignored."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return concatenate([self])
        device = self.device if len(self._tensors) else None
        payload = PandasTensorCollection(self.infos, **{k: t.cpu() for k, t in self._tensors.items()})
        gathered = [None] * dist.get_world_size()
        dist.all_gather_object(gathered, payload)
        out = concatenate(gathered)
        return out.to(device) if device is not None and len(out.tensors) else out

    def __getstate__(self):
        state = super().__getstate__()
        state['infos'] = self.infos
        state['meta'] = self.meta
        return state

    def __setstate__(self, state):
        PandasTensorCollection.__init__(self, state['infos'], **state['tensors'])
        self.meta = state['meta']


def concatenate(datas):
    datas = [d for d in datas if len(d) > 0]
    if not datas:
        return PandasTensorCollection(infos=pd.DataFrame())
    assert all(type(d) is type(datas[0]) for d in datas)
    if len(datas) == 1:     # a fresh collection over the same rows (its constructor re-indexes a copy of infos)
        return PandasTensorCollection(infos=datas[0].infos, **dict(datas[0].tensors))
    infos = pd.concat([d.infos for d in datas], axis=0, sort=False).reset_index(drop=True)
    tensors = {k: torch.cat([getattr(d, k) for d in datas], dim=0) for k in datas[0].tensors}
    return PandasTensorCollection(infos=infos, **tensors)
