"""Validation loop, meters and loaders (mirror of util.py:11-133).  `get_imagenet_validation`
keeps the reference's signature and transforms (util.py:11-36) and is used when `args.val_dir`
holds `imagenet/val`; where that dataset does not exist (this container, the GPU box)
it falls back to `synthetic_loader`, which produces batches of the same shape and dtype."""
import os
import time

import torch


class SyntheticImages(torch.utils.data.Dataset):
    """`n` seeded 3x224x224 fp32 images with random labels; `.targets` like ImageFolder."""

    def __init__(self, n, num_classes=1000, size=224, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.targets = torch.randint(num_classes, (n,), generator=g).tolist()
        self.size, self.seed, self.n = size, seed, n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1000003 + i)
        return torch.randn(3, self.size, self.size, generator=g), self.targets[i]


def synthetic_loader(n_images, batch_size, size=224, seed=0, workers=0):
    return torch.utils.data.DataLoader(SyntheticImages(n_images, size=size, seed=seed),
                                       batch_size=batch_size, shuffle=False, num_workers=workers,
                                       pin_memory=True)


def get_imagenet_validation(args):
    """ImageNet validation loader (util.py:11-36): Resize(256) + CenterCrop(224) + Normalize, or the
    EfficientNet image size with bicubic resize for `efficientnet_*` (resolved lazily: the third-party
    package is optional).  Without `<val_dir>/imagenet/val` a synthetic loader of the same shape is
    returned (`args.images` images if given, else 256)."""
    val_dir = getattr(args, 'val_dir', None)
    root = os.path.join(val_dir, 'imagenet', 'val') if val_dir else None
    image_size, bicubic = 224, False
    if 'efficientnet' in getattr(args, 'arch', ''):
        try:
            from efficientnet_pytorch import EfficientNet
        except ImportError as e:
            raise ImportError("efficientnet_* needs the third-party efficientnet_pytorch package (a dependency of the "
                              "reference, util.py:4; not in this image)") from e
        image_size, bicubic = EfficientNet.get_image_size(args.arch.replace('_', '-')), True
    if root is None or not os.path.isdir(root):
        return synthetic_loader(getattr(args, 'images', None) or 256, args.batch_size, size=image_size,
                                workers=getattr(args, 'workers', 0))
    import PIL
    import torchvision.datasets as datasets
    import torchvision.transforms as transforms
    normalize = transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    resize = (transforms.Resize(image_size, interpolation=PIL.Image.BICUBIC) if bicubic
              else transforms.Resize(256))
    tf = transforms.Compose([resize, transforms.CenterCrop(image_size), transforms.ToTensor(), normalize])
    return torch.utils.data.DataLoader(datasets.ImageFolder(root, tf), batch_size=args.batch_size,
                                       shuffle=False, num_workers=getattr(args, 'workers', 0),
                                       pin_memory=True)


def accuracy(output, target, topk=1):
    """Top-k accuracy in percent."""
    with torch.no_grad():
        pred = output.topk(topk, 1, True, True)[1]
        hit = pred.eq(target.view(-1, 1)).any(dim=1).float().sum()
        return (hit * (100.0 / target.size(0))).item()


class AverageMeter:
    def __init__(self, name, fmt=':f'):
        self.name, self.fmt = name, fmt
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

    def __str__(self):
        return ('{name} {val' + self.fmt + '} ({avg' + self.fmt + '})').format(**self.__dict__)


class ProgressMeter:
    def __init__(self, num_batches, meters, prefix=""):
        width = len(str(num_batches))
        self.fmt = '[{:' + str(width) + 'd}/' + str(num_batches) + ']'
        self.meters, self.prefix = meters, prefix

    def display(self, batch):
        print('\t'.join([self.prefix + self.fmt.format(batch)] + [str(m) for m in self.meters]))


def validate(val_loader, model, criterion, args, verbose=True, pct=1.0):
    """Evaluate `model` on the first `pct` of the loader; returns (avg loss, top-1 %)."""
    batch_time = AverageMeter('Time', ':6.3f')
    losses = AverageMeter('Loss', ':.4e')
    top1 = AverageMeter('Acc@1', ':6.2f')
    progress = ProgressMeter(len(val_loader), [batch_time, losses, top1], prefix='Test: ')
    model.eval()
    eval_samples = round(pct * len(val_loader.dataset.targets))
    seen = 0
    gpu = getattr(args, 'gpu', None)
    with torch.no_grad():
        end = time.time()
        for i, (images, target) in enumerate(val_loader):
            if gpu is not None:
                images = images.cuda(gpu, non_blocking=True)
            target = target.cuda(gpu, non_blocking=True)
            seen += len(target)
            output = model(images)
            loss = criterion(output, target)
            losses.update(loss.item(), images.size(0))
            top1.update(accuracy(output, target, topk=1), images.size(0))
            batch_time.update(time.time() - end)
            end = time.time()
            if verbose and i % getattr(args, 'print_freq', 10) == 0:
                progress.display(i)
            if seen >= eval_samples:
                break
    if verbose:
        print(' * Acc@1 {:.3f} '.format(top1.avg))
    return losses.avg, top1.avg
