"""Wire messages of the fan-out path, built at run time with the protobuf runtime (no protoc in the image):
channeldpb.Packet / MessagePack / ChannelDataUpdateMessage (pkg/channeldpb/channeld.proto:10-34,333-340) and a channel data
message shaped like internal/testpb/test.proto's TestChannelDataMessage (text, num) extended with a repeated and a map field so
that every merge rule (last scalar wins, lists append, map entries override) is exercised.  Test infrastructure."""
from google.protobuf import any_pb2, descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto


def _field(m, name, num, typ, label=_F.LABEL_OPTIONAL, type_name=None):
    f = m.field.add()
    f.name, f.number, f.type, f.label = name, num, typ, label
    if type_name:
        f.type_name = type_name
    return f


def _build():
    pool = descriptor_pool.Default()
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "chd_test_wire.proto", "chdtest", "proto3"
    fd.dependency.append("google/protobuf/any.proto")
    mp = fd.message_type.add()
    mp.name = "MessagePack"
    _field(mp, "channelId", 1, _F.TYPE_UINT32)
    _field(mp, "broadcast", 2, _F.TYPE_UINT32)
    _field(mp, "stubId", 3, _F.TYPE_UINT32)
    _field(mp, "msgType", 4, _F.TYPE_UINT32)
    _field(mp, "msgBody", 5, _F.TYPE_BYTES)
    pk = fd.message_type.add()
    pk.name = "Packet"
    _field(pk, "messages", 1, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, ".chdtest.MessagePack")
    cdu = fd.message_type.add()
    cdu.name = "ChannelDataUpdateMessage"
    _field(cdu, "data", 1, _F.TYPE_MESSAGE, type_name=".google.protobuf.Any")
    _field(cdu, "contextConnId", 2, _F.TYPE_UINT32)
    d = fd.message_type.add()
    d.name = "TestChannelDataMessage"
    _field(d, "text", 1, _F.TYPE_STRING)
    _field(d, "num", 2, _F.TYPE_UINT32)
    _field(d, "list", 3, _F.TYPE_STRING, _F.LABEL_REPEATED)
    kv = d.nested_type.add()
    kv.name = "KvEntry"
    kv.options.map_entry = True
    _field(kv, "key", 1, _F.TYPE_UINT32)
    _field(kv, "value", 2, _F.TYPE_STRING)
    _field(d, "kv", 4, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, ".chdtest.TestChannelDataMessage.KvEntry")
    any_pb2.Any()  # make sure any.proto is in the default pool
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("chdtest." + n))  # noqa: E731
    return get("Packet"), get("MessagePack"), get("ChannelDataUpdateMessage"), get("TestChannelDataMessage")


Packet, MessagePack, ChannelDataUpdateMessage, TestChannelDataMessage = _build()
TYPE_URL = "type.googleapis.com/chdtest.TestChannelDataMessage"
MSG_CHANNEL_DATA_UPDATE = 8  # channeld.proto:121
