/* hd_yaml.c -- compact parser for the YAML subset hypredrive accepts (block mappings, scalars,
 * one-line flow mappings, "- " sequence items, include:, "--a:b:c value" overrides).
 * Re-implementation of the behaviour documented for the reference's src/internal/yaml.c
 * (line grammar :1013-1205, value normalisation :2507-2531, overrides :2188-2330):
 *   - everything after the first '#' on a line is dropped; blank lines are ignored;
 *   - indentation / base_indent (auto-detected) = nesting level; tabs, inconsistent indents
 *     and level jumps raise the ERROR_YAML_* bits;
 *   - the first ':' splits key and value; values are trimmed, unquoted and lower-cased unless
 *     the key contains "name" (file names keep their case). */
#include "hd_internal.h"
#include <ctype.h>
#include <stdlib.h>
#include <string.h>

static char *dup_range(const char *s, size_t n)
{
   char *r = malloc(n + 1);
   memcpy(r, s, n);
   r[n] = 0;
   return r;
}

static char *trim_dup(const char *s, size_t n)
{
   while (n && isspace((unsigned char)*s)) { s++; n--; }
   while (n && isspace((unsigned char)s[n - 1])) n--;
   if (n >= 2 && ((s[0] == '"' && s[n - 1] == '"') || (s[0] == '\'' && s[n - 1] == '\''))) { s++; n -= 2; }
   return dup_range(s, n);
}

static hd_node *node_new(const char *key, const char *rawval, int level)
{
   hd_node *n = calloc(1, sizeof(hd_node));
   n->key     = strdup(key ? key : "");
   n->raw_val = strdup(rawval ? rawval : "");
   n->val     = strdup(n->raw_val);
   if (!strstr(n->key, "name"))
      for (char *p = n->val; *p; p++) *p = (char)tolower((unsigned char)*p);
   n->level = level;
   return n;
}

static void node_append(hd_node *parent, hd_node *c)
{
   c->parent = parent;
   c->next   = NULL;
   if (!parent->child) { parent->child = c; return; }
   hd_node *t = parent->child;
   while (t->next) t = t->next;
   t->next = c;
}

void hd_yaml_free(hd_node *n)
{
   while (n)
   {
      hd_node *nx = n->next;
      hd_yaml_free(n->child);
      free(n->key); free(n->val); free(n->raw_val);
      free(n);
      n = nx;
   }
}

hd_node *hd_yaml_find(hd_node *parent, const char *key)
{
   if (!parent) return NULL;
   for (hd_node *c = parent->child; c; c = c->next)
      if (!strcmp(c->key, key)) return c;
   return NULL;
}

/* "{a: b, c: {d: e}}" -> children of `parent` */
static const char *parse_flow(hd_node *parent, const char *s, int level)
{
   /* s points just after '{' */
   while (*s)
   {
      while (*s && (isspace((unsigned char)*s) || *s == ',')) s++;
      if (*s == '}') return s + 1;
      const char *k0 = s;
      while (*s && *s != ':' && *s != '}' && *s != ',') s++;
      if (*s != ':') { hd_err_set(HYPREDRV_ERROR_YAML_INVALID_DIVISOR); hd_err_msg("flow mapping entry without ':' near '%.20s'", k0); return s; }
      char *key = trim_dup(k0, (size_t)(s - k0));
      s++;
      while (*s && isspace((unsigned char)*s)) s++;
      if (*s == '{')
      {
         hd_node *n = node_new(key, "", level);
         node_append(parent, n);
         s = parse_flow(n, s + 1, level + 1);
      }
      else
      {
         const char *v0 = s;
         while (*s && *s != ',' && *s != '}') s++;
         char *val = trim_dup(v0, (size_t)(s - v0));
         node_append(parent, node_new(key, val, level));
         free(val);
      }
      free(key);
   }
   return s;
}

static char *read_file(const char *path)
{
   FILE *fp = fopen(path, "rb");
   if (!fp) return NULL;
   fseek(fp, 0, SEEK_END);
   long sz = ftell(fp);
   fseek(fp, 0, SEEK_SET);
   char *buf = malloc((size_t)sz + 1);
   size_t rd = fread(buf, 1, (size_t)sz, fp);
   buf[rd] = 0;
   fclose(fp);
   return buf;
}

static int parse_into(hd_node *root, const char *text, const char *base_dir, int depth);

static void splice_include(hd_node *parent, const char *fname, const char *base_dir, int depth)
{
   char path[2048];
   if (fname[0] == '/') { hd_err_set(HYPREDRV_ERROR_FILE_NOT_FOUND); hd_err_msg("absolute include path rejected: %s", fname); return; }
   snprintf(path, sizeof(path), "%s%s%s", base_dir ? base_dir : ".", "/", fname);
   char *txt = read_file(path);
   if (!txt) { hd_err_set(HYPREDRV_ERROR_FILE_NOT_FOUND); hd_err_msg("cannot open include file: %s", path); return; }
   hd_node *sub = calloc(1, sizeof(hd_node));
   sub->key = strdup(""); sub->val = strdup(""); sub->raw_val = strdup(""); sub->level = -1;
   char dir[2048];
   snprintf(dir, sizeof(dir), "%s", path);
   char *slash = strrchr(dir, '/');
   if (slash) *slash = 0;
   parse_into(sub, txt, dir, depth + 1);
   free(txt);
   /* move the children of `sub` under `parent` with adjusted levels */
   hd_node *c = sub->child;
   sub->child = NULL;
   while (c)
   {
      hd_node *nx = c->next;
      node_append(parent, c);
      c = nx;
   }
   hd_yaml_free(sub);
}

static int parse_into(hd_node *root, const char *text, const char *base_dir, int depth)
{
   if (depth > 8) { hd_err_set(HYPREDRV_ERROR_YAML_TREE_INVALID); hd_err_msg("include nesting too deep"); return 1; }
   hd_node *stack[64];
   int      stack_indent[64];
   int      sp = 0, base = 0;
   stack[0] = root; stack_indent[0] = -1;
   const char *p = text;
   int         lineno = 0;
   while (*p)
   {
      const char *eol = strchr(p, '\n');
      size_t      len = eol ? (size_t)(eol - p) : strlen(p);
      lineno++;
      char *line = dup_range(p, len);
      p += len + (eol ? 1 : 0);
      char *hash = strchr(line, '#');
      if (hash) *hash = 0;
      size_t n = strlen(line);
      while (n && isspace((unsigned char)line[n - 1])) line[--n] = 0;
      /* indentation */
      int indent = 0, tabs = 0, spaces = 0;
      while (line[indent] == ' ' || line[indent] == '\t') { if (line[indent] == '\t') tabs++; else spaces++; indent++; }
      if (line[indent] == 0) { free(line); continue; }
      if (tabs)
      {
         hd_err_set(spaces ? HYPREDRV_ERROR_YAML_MIXED_INDENT : HYPREDRV_ERROR_YAML_INVALID_INDENT);
         hd_err_msg("line %d: tab characters in indentation", lineno);
         free(line);
         return 1;
      }
      char *body = line + indent;
      int   is_seq = 0;
      if (body[0] == '-' && (body[1] == ' ' || body[1] == 0)) { is_seq = 1; }
      if (indent > 0 && base == 0) base = indent;
      if (base > 0 && indent % base != 0 && !(indent % 2 == 0 && sp > 0 && stack[sp]->is_seq_item))
      {
         hd_err_set(HYPREDRV_ERROR_YAML_INCONSISTENT_INDENT);
         hd_err_msg("line %d: indentation %d is not a multiple of %d", lineno, indent, base);
         free(line);
         return 1;
      }
      /* pop to the parent whose indent is smaller */
      while (sp > 0 && stack_indent[sp] >= indent) sp--;
      if (sp > 0 || indent > 0)
      {
         int parent_indent = stack_indent[sp];
         if (base > 0 && parent_indent >= 0 && indent - parent_indent > base && !stack[sp]->is_seq_item)
         {
            hd_err_set(HYPREDRV_ERROR_YAML_INVALID_INDENT_JUMP);
            hd_err_msg("line %d: indentation jumps more than one level", lineno);
            free(line);
            return 1;
         }
         if (sp == 0 && indent > 0 && root->level >= -1 && root->child == NULL)
         {
            hd_err_set(HYPREDRV_ERROR_YAML_INVALID_BASE_INDENT);
            hd_err_msg("line %d: first entry must not be indented", lineno);
            free(line);
            return 1;
         }
      }
      hd_node *parent = stack[sp];
      int      level  = parent->level + 1;
      if (is_seq)
      {
         hd_node *item = node_new("-", "", level);
         item->is_seq_item = 1;
         node_append(parent, item);
         if (sp + 1 < 63) { sp++; stack[sp] = item; stack_indent[sp] = indent; }
         body += 1;
         while (*body == ' ') { body++; }
         if (*body == 0) { free(line); continue; }
         indent = (int)(body - line);
         parent = item;
         level  = item->level + 1;
         if (*body == '{') { parse_flow(item, body + 1, level); free(line); continue; }
      }
      char *colon = strchr(body, ':');
      if (!colon)
      {
         if (is_seq) { free(parent->val); free(parent->raw_val); parent->raw_val = trim_dup(body, strlen(body)); parent->val = strdup(parent->raw_val); free(line); continue; }
         hd_err_set(HYPREDRV_ERROR_YAML_INVALID_DIVISOR);
         hd_err_msg("line %d: missing ':' in '%s'", lineno, body);
         free(line);
         return 1;
      }
      char *key = trim_dup(body, (size_t)(colon - body));
      char *val = trim_dup(colon + 1, strlen(colon + 1));
      if (!strcmp(key, "include") && val[0])
      {
         splice_include(parent, val, base_dir, depth);
      }
      else if (val[0] == '{')
      {
         hd_node *nn = node_new(key, "", level);
         node_append(parent, nn);
         parse_flow(nn, val + 1, level + 1);
      }
      else
      {
         hd_node *nn = node_new(key, val, level);
         node_append(parent, nn);
         if (sp + 1 < 63) { sp++; stack[sp] = nn; stack_indent[sp] = indent; }
      }
      free(key); free(val); free(line);
   }
   return 0;
}

/* `include:` followed by a sequence of file names (reference examples/ex8-multi-1.yml: one
 * preconditioner variant per file): every listed file becomes one sequence item of the parent,
 * holding the file's tree.  The scalar form `include: file.yml` is spliced while parsing. */
static void expand_include_lists(hd_node *n, const char *base_dir)
{
   for (hd_node *c = n->child; c; c = c->next) expand_include_lists(c, base_dir);
   hd_node *prev = NULL, *c = n->child;
   while (c)
   {
      hd_node *nx = c->next;
      if (!strcmp(c->key, "include") && !c->val[0] && c->child && c->child->is_seq_item)
      {
         /* unlink the include node, append one item per file at the end of the parent */
         if (prev) prev->next = nx; else n->child = nx;
         c->next = NULL;
         for (hd_node *it = c->child; it; it = it->next)
         {
            hd_node *item = node_new("-", "", n->level + 1);
            item->is_seq_item = 1;
            node_append(n, item);
            splice_include(item, it->raw_val, base_dir, 1);
         }
         hd_yaml_free(c);
         c = nx;
         continue;
      }
      prev = c;
      c    = nx;
   }
}

hd_node *hd_yaml_parse(const char *text, const char *base_dir)
{
   hd_node *root = calloc(1, sizeof(hd_node));
   root->key = strdup(""); root->val = strdup(""); root->raw_val = strdup(""); root->level = -1;
   uint32_t before = hd_err_get();
   if (!text) { hd_err_set(HYPREDRV_ERROR_YAML_TREE_NULL); hd_yaml_free(root); return NULL; }
   parse_into(root, text, base_dir, 0);
   if (!(hd_err_get() & ~before)) expand_include_lists(root, base_dir);
   if (hd_err_get() & ~before) { hd_yaml_free(root); return NULL; }
   return root;
}

/* "--a:b:c" value : create or overwrite the node at that path */
int hd_yaml_override(hd_node *root, const char *path, const char *value)
{
   while (*path == '-') path++;
   if (!*path) return 1;
   hd_node *cur = root;
   char    *tmp = strdup(path);
   char    *save = NULL;
   for (char *tok = strtok_r(tmp, ":", &save); tok; tok = strtok_r(NULL, ":", &save))
   {
      hd_node *c = hd_yaml_find(cur, tok);
      if (!c)
      {
         c = node_new(tok, "", cur->level + 1);
         node_append(cur, c);
      }
      /* a value-only parent ("solver: pcg") becomes a block when overridden below it */
      if (cur != root && cur->val[0] && !hd_yaml_find(cur, cur->val) && strcmp(cur->val, tok))
      {
         /* keep the scalar: "solver: pcg" + "--solver:pcg:max_iter" -> block */
      }
      cur = c;
   }
   free(cur->val); free(cur->raw_val);
   cur->raw_val = strdup(value);
   cur->val     = strdup(value);
   if (!strstr(cur->key, "name"))
      for (char *p = cur->val; *p; p++) *p = (char)tolower((unsigned char)*p);
   /* ancestors that were scalars naming this child become plain blocks */
   for (hd_node *a = cur->parent; a && a != root; a = a->parent)
      if (a->val[0]) { a->val[0] = 0; a->raw_val[0] = 0; }
   free(tmp);
   return 0;
}

void hd_yaml_print(const hd_node *n, FILE *fp)
{
   for (const hd_node *c = n ? n->child : NULL; c; c = c->next)
   {
      const char *flag = c->invalid == 1 ? "   <-- invalid key" : (c->invalid == 2 ? "   <-- invalid value" : "");
      fprintf(fp, "%*s%s: %s%s\n", 2 * (c->level > 0 ? c->level : 0), "", c->key, c->raw_val, flag);
      hd_yaml_print(c, fp);
   }
}
