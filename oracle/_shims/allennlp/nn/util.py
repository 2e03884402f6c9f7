"""Shim for the three allennlp.nn.util functions the reference imports
(src/models/lipreader/better_model.py:3).  allennlp is un-vendored and un-pinned in the reference;
semantics restated from allennlp 0.7/0.8 (Dec 2018).  Used ONLY to import the reference as an
oracle in the build container (oracle/ref_harness.py)."""
import torch


def masked_log_softmax(vector, mask, dim=-1):
    if mask is not None:
        mask = mask.float()
        while mask.dim() < vector.dim():
            mask = mask.unsqueeze(1)
        vector = vector + (mask + 1e-45).log()
    return torch.nn.functional.log_softmax(vector, dim=dim)


def masked_softmax(vector, mask, dim=-1):
    if mask is None:
        return torch.nn.functional.softmax(vector, dim=dim)
    mask = mask.float()
    while mask.dim() < vector.dim():
        mask = mask.unsqueeze(1)
    result = torch.nn.functional.softmax(vector * mask, dim=dim)
    result = result * mask
    return result / (result.sum(dim=dim, keepdim=True) + 1e-13)


def sort_batch_by_length(tensor, sequence_lengths):
    sorted_lengths, permutation = sequence_lengths.sort(0, descending=True)
    sorted_tensor = tensor.index_select(0, permutation)
    index_range = torch.arange(0, len(sequence_lengths), device=sequence_lengths.device)
    _, reverse_mapping = permutation.sort(0, descending=False)
    restoration = index_range.index_select(0, reverse_mapping)
    return sorted_tensor, sorted_lengths, restoration, permutation
