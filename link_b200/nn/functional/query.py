import torch

from link_b200 import _capi

__all__ = ['sphashquery', 'HashTable']


class HashTable:
    """Device hash table over int64 keys -> row index (lk_table_build)."""

    def __init__(self, references: torch.Tensor):
        references = references.contiguous()
        assert references.dtype == torch.int64 and references.ndim == 1
        L = _capi.lib()
        self.n = references.shape[0]
        self.capacity = int(L.lk_table_capacity(self.n))
        self.table = torch.empty(self.capacity * 16, dtype=torch.uint8, device=references.device)
        _capi.check(L.lk_table_build(_capi.ptr(references), self.n, _capi.ptr(self.table),
                                     self.capacity, _capi.stream()), 'lk_table_build')

    def query(self, queries: torch.Tensor) -> torch.Tensor:
        sizes = queries.size()
        q = queries.contiguous().view(-1)
        assert q.dtype == torch.int64
        out = torch.empty(q.shape[0], dtype=torch.int64, device=q.device)
        _capi.check(_capi.lib().lk_table_query(_capi.ptr(q), q.shape[0], _capi.ptr(self.table),
                                               self.capacity, _capi.ptr(out), _capi.stream()),
                    'lk_table_query')
        return out.view(*sizes)


def sphashquery(queries: torch.Tensor, references: torch.Tensor) -> torch.Tensor:
    """Reference `F.sphashquery` (torchsparse/nn/functional/query.py:8-33): for every query hash
    the index of the equal reference hash, or -1.  Same shape as `queries`, int64."""
    return HashTable(references).query(queries)
