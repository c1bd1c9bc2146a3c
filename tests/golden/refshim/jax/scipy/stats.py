import math
import torch


class norm:
    @staticmethod
    def pdf(x, loc=0., scale=1.):
        z = (x - loc) / scale
        return torch.exp(-0.5 * z * z) / (scale * math.sqrt(2 * math.pi))

    @staticmethod
    def logpdf(x, loc=0., scale=1.):
        z = (x - loc) / scale
        return -0.5 * z * z - torch.log(torch.as_tensor(scale * math.sqrt(2 * math.pi)))
