"""Device discovery with the call shape of the reference's helper.device_list()
(src/_helper/helper.py:1172-1224): a list whose items can be passed to NUFFT(device)."""


def device_list():
    import torch
    return [torch.device('cuda', i) for i in range(torch.cuda.device_count())]


def diagnose():
    import torch
    for i in range(torch.cuda.device_count()):
        p = torch.cuda.get_device_properties(i)
        print('cuda:%d %s sm_%d%d %.0f GB' % (i, p.name, p.major, p.minor, p.total_memory / 2**30))
