"""Module path helpers used by the runner -- mirror of the parts of
src/myutils/pytorch/module_util.py the hot path touches (:16-27, :34-57)."""
from torch.nn import DataParallel
from torch.nn.parallel import DistributedDataParallel


def freeze_module_params(module):
    for param in module.parameters():
        param.requires_grad = False


def unfreeze_module_params(module):
    for param in module.parameters():
        param.requires_grad = True


def get_updatable_param_names(module):
    return [name for name, param in module.named_parameters() if param.requires_grad]


def get_module(root_module, module_path):
    """module_util.py:34-57: dotted path lookup; prints and returns None on a bad path."""
    module_names = module_path.split('.')
    module = root_module
    for module_name in module_names:
        if not hasattr(module, module_name):
            if isinstance(module, (DataParallel, DistributedDataParallel)):
                module = module.module
                if not hasattr(module, module_name):
                    print('`{}` of `{}` could not be reached in `{}`'.format(module_name, module_path,
                                                                             type(root_module).__name__))
                    return None
            else:
                print('`{}` of `{}` could not be reached in `{}`'.format(module_name, module_path,
                                                                         type(root_module).__name__))
                return None
        module = getattr(module, module_name)
    return module
