"""The PyTorch C++ extension (csrc/torch_binding.cpp -> ntm_b200_torch.so) loads without a GPU and registers torch.ops.ntm.*
with the documented schemas; shape / device validation raises RuntimeError (TORCH_CHECK) before any compute."""
import pytest
import torch

import ntm_b200
from ntm_b200 import lib

OPS = ["prepare", "destroy", "gru_forward", "gru_forward_out", "diffdel_forward", "delay_forward", "delay_check", "esr_sums",
       "esr_sums_rows"]


def test_ops_are_registered_with_schemas():
    ops = lib.ops()
    for name in OPS:
        assert hasattr(ops, name), name
    schema = str(torch.ops.ntm.gru_forward_out.default._schema)
    assert "Tensor(a!) y" in schema and "Tensor(b!) h_out" in schema          # declared as writing its output arguments
    assert "Tensor? h_in" in str(torch.ops.ntm.gru_forward.default._schema)


def test_ops_validate_before_compute():
    ops = lib.ops()
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.gru_forward(0, 0, torch.zeros(1, 1, 4), None, False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.delay_forward(torch.zeros(1, 1, 4), torch.zeros(1, 1, 4), torch.zeros(1, 1, 2), False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.esr_sums(torch.zeros(1, 1, 4), torch.zeros(1, 1, 4), True)
    w = torch.zeros(192, 8)
    with pytest.raises(RuntimeError):                                          # hidden size 8: not built (or no device)
        ops.prepare(torch.zeros(24, 1), torch.zeros(24, 8), torch.zeros(24), torch.zeros(24), torch.zeros(1, 8), None, 0)
    del w


def test_unsupported_hidden_size_is_a_clear_error():
    m = ntm_b200.RNN()                       # the reference's default hidden_size=8 (code/model.py:22)
    assert m.hidden_size == 8
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 1, 8))
