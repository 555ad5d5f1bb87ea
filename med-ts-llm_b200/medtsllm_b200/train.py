"""Autograd (training) path: forward with activation stash + backward through the frozen backbone
for the adapter gradients.  See DESIGN.md §training."""
from ._lib import MtsError


def forward_train(model, inputs):
    raise MtsError("the training path (adapter gradients through the frozen backbone) is not built yet; "
                   "wrap inference calls in torch.no_grad()")
