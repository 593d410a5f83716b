#include "xml_dom.h"

#include <cstdio>
#include <cstring>

namespace rdc {
namespace {

struct Cursor {
  const char* p;
  const char* end;
  size_t line = 1;

  bool eof() const { return p >= end; }
  char peek() const { return p < end ? *p : '\0'; }
  void advance() {
    if (p < end) {
      if (*p == '\n') ++line;
      ++p;
    }
  }
  bool starts_with(const char* s) const {
    size_t n = std::strlen(s);
    return (size_t)(end - p) >= n && std::memcmp(p, s, n) == 0;
  }
  void skip(size_t n) {
    while (n-- && p < end) advance();
  }
  [[noreturn]] void fail(const std::string& what) const {
    throw XmlError("xml: " + what + " at line " + std::to_string(line));
  }
};

bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
bool is_name_char(char c) { return !(is_space(c) || c == '/' || c == '>' || c == '=' || c == '<' || c == '\0'); }

void skip_space(Cursor& c) {
  while (!c.eof() && is_space(c.peek())) c.advance();
}

void skip_until(Cursor& c, const char* terminator) {
  while (!c.eof() && !c.starts_with(terminator)) c.advance();
  if (c.eof()) c.fail(std::string("unterminated construct, expected ") + terminator);
  c.skip(std::strlen(terminator));
}

// <!DOCTYPE ...> may nest an internal subset in [ ].
void skip_doctype(Cursor& c) {
  int depth = 0;
  while (!c.eof()) {
    char ch = c.peek();
    if (ch == '[') ++depth;
    else if (ch == ']') --depth;
    else if (ch == '>' && depth <= 0) {
      c.advance();
      return;
    }
    c.advance();
  }
  c.fail("unterminated DOCTYPE");
}

void append_utf8(std::string& out, unsigned long cp) {
  if (cp < 0x80) out += (char)cp;
  else if (cp < 0x800) {
    out += (char)(0xC0 | (cp >> 6));
    out += (char)(0x80 | (cp & 0x3F));
  } else if (cp < 0x10000) {
    out += (char)(0xE0 | (cp >> 12));
    out += (char)(0x80 | ((cp >> 6) & 0x3F));
    out += (char)(0x80 | (cp & 0x3F));
  } else {
    out += (char)(0xF0 | (cp >> 18));
    out += (char)(0x80 | ((cp >> 12) & 0x3F));
    out += (char)(0x80 | ((cp >> 6) & 0x3F));
    out += (char)(0x80 | (cp & 0x3F));
  }
}

std::string decode_entities(const char* b, const char* e) {
  std::string out;
  out.reserve(e - b);
  while (b < e) {
    if (*b != '&') {
      out += *b++;
      continue;
    }
    const char* semi = (const char*)std::memchr(b, ';', e - b);
    if (!semi) {
      out += *b++;
      continue;
    }
    std::string ent(b + 1, semi);
    if (ent == "amp") out += '&';
    else if (ent == "lt") out += '<';
    else if (ent == "gt") out += '>';
    else if (ent == "quot") out += '"';
    else if (ent == "apos") out += '\'';
    else if (ent.size() > 1 && ent[0] == '#') {
      unsigned long cp = (ent[1] == 'x' || ent[1] == 'X') ? std::strtoul(ent.c_str() + 2, nullptr, 16)
                                                            : std::strtoul(ent.c_str() + 1, nullptr, 10);
      append_utf8(out, cp);
    } else {
      out.append(b, semi + 1);  // unknown entity: keep verbatim
    }
    b = semi + 1;
  }
  return out;
}

std::string parse_name(Cursor& c) {
  const char* b = c.p;
  while (!c.eof() && is_name_char(c.peek())) c.advance();
  if (c.p == b) c.fail("expected a name");
  return std::string(b, c.p);
}

// Skips everything that is not an element start. Returns false at end of input or at a closing tag.
bool seek_element(Cursor& c) {
  for (;;) {
    while (!c.eof() && c.peek() != '<') c.advance();  // character data
    if (c.eof()) return false;
    if (c.starts_with("<!--")) skip_until(c, "-->");
    else if (c.starts_with("<![CDATA[")) skip_until(c, "]]>");
    else if (c.starts_with("<!")) skip_doctype(c);
    else if (c.starts_with("<?")) skip_until(c, "?>");
    else if (c.starts_with("</")) return false;
    else return true;
  }
}

std::unique_ptr<XmlElement> parse_element(Cursor& c) {
  c.advance();  // '<'
  auto el = std::make_unique<XmlElement>();
  el->name = parse_name(c);
  for (;;) {
    skip_space(c);
    if (c.eof()) c.fail("unterminated tag <" + el->name);
    if (c.peek() == '/') {
      c.advance();
      if (c.peek() != '>') c.fail("expected '>' after '/'");
      c.advance();
      return el;
    }
    if (c.peek() == '>') {
      c.advance();
      break;
    }
    std::string key = parse_name(c);
    skip_space(c);
    if (c.peek() != '=') c.fail("expected '=' after attribute " + key);
    c.advance();
    skip_space(c);
    char quote = c.peek();
    if (quote != '"' && quote != '\'') c.fail("expected a quoted value for attribute " + key);
    c.advance();
    const char* b = c.p;
    while (!c.eof() && c.peek() != quote) c.advance();
    if (c.eof()) c.fail("unterminated value of attribute " + key);
    el->attrs.emplace_back(std::move(key), decode_entities(b, c.p));
    c.advance();
  }
  // content
  while (seek_element(c)) el->children.push_back(parse_element(c));
  if (c.eof()) c.fail("missing </" + el->name + ">");
  c.skip(2);  // "</"
  std::string closing = parse_name(c);
  if (closing != el->name) c.fail("</" + closing + "> closes <" + el->name + ">");
  skip_space(c);
  if (c.peek() != '>') c.fail("expected '>' in closing tag");
  c.advance();
  return el;
}

}  // namespace

std::unique_ptr<XmlElement> xml_parse(const char* text, size_t len) {
  Cursor c{text, text + len};
  if (len >= 3 && (unsigned char)text[0] == 0xEF && (unsigned char)text[1] == 0xBB && (unsigned char)text[2] == 0xBF)
    c.p += 3;  // UTF-8 byte order mark
  if (!seek_element(c)) throw XmlError("xml: document has no root element");
  return parse_element(c);
}

std::unique_ptr<XmlElement> xml_parse_file(const std::string& path) {
  std::FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw XmlError("xml: cannot open " + path);
  std::string data;
  char buf[1 << 16];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, n);
  std::fclose(f);
  return xml_parse(data.data(), data.size());
}

void xml_dump(const XmlElement& e, int depth, std::string& out) {
  out.append((size_t)depth, ' ');
  out += e.name;
  for (auto& a : e.attrs) {
    out += ' ';
    out += a.first;
    out += '=';
    out += a.second;
  }
  out += '\n';
  for (auto& c : e.children) xml_dump(*c, depth + 1, out);
}

}  // namespace rdc
