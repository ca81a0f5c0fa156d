"""Image pre-processing (mirror of /root/reference/data/helper.py:9-27): ToTensor + ImageNet normalisation, with a
Resize(224) only for img_size == 224."""
import torchvision.transforms as transforms

_MEAN, _STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def get_transforms(args):
    steps = []
    if args.img_size == 224:
        steps.append(transforms.Resize(224))
    elif args.img_size != 512:
        raise ValueError("img_size must be 224 or 512 (reference data/helper.py)")
    steps += [transforms.ToTensor(), transforms.Normalize(_MEAN, _STD)]
    return transforms.Compose(steps)
