/* shim: header included but unused by the reference TUs built into oracle/_ref */
#pragma once
