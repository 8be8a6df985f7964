def get_external_ip():
    return "127.0.0.1"


def ip_to_int(ip):
    return 0


def ip_version(ip):
    return 4
