// Stand-in for libProtoBuf/protobuf_aux.hpp -- TEST INFRASTRUCTURE: file I/O of messages is outside what oracle/_ref pins.
#pragma once
#include <QString>
#include <cstdlib>
template <class M> void write_message_binary(QString, const M &) {}
template <class M> bool parse_message_binary(QString, M &) { abort(); }
template <class M> void parse_message_from_text_file(QString, M &) { abort(); }
