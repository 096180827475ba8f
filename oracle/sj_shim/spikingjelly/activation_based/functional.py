def reset_net(net):
    for m in net.modules():
        if hasattr(m, "reset"):
            m.reset()
